// Internal state of libiris shared by the extern "C" translation units (not a public header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/iris.h"
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

std::string& last_error();   // thread-local text behind iris_last_error()

inline int fail(int code, const std::string& msg) {
    last_error() = msg;
    return code;
}
inline int cuda_fail(cudaError_t e, const char* what) {
    last_error() = std::string(what) + ": " + cudaGetErrorString(e);
    return IRIS_ERR_CUDA;
}
#define CU(x)                                               \
    do {                                                    \
        cudaError_t e_ = (x);                               \
        if (e_ != cudaSuccess) return ::iris::cuda_fail(e_, #x); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct Bank {
    bool ready = false;
    bool spec = false;                 // items are pre-computed spectrograms [257, t, 2C] (k_spec.cu)
    int n_items = 0, n_chan = 0, n_classes = 0;
    std::vector<int64_t> offsets;      // samples per channel, cumulative
    std::vector<int64_t> pad_offsets;  // padded floats per channel, cumulative
    std::vector<int32_t> n_frames;
    int max_frames = 0;
    DevBuf padded, activity, labels, d_n_frames;
    std::vector<uint8_t> h_activity;   // host mirror (voice bank)
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace iris

struct iris_ctx {
    using DevBuf = iris::DevBuf;
    using Bank = iris::Bank;
    using Seg = iris::Seg;
    int device = 0;
    int num_sms = 148;
    Bank banks[3];
    DevBuf tw, ts, whalf;
    // mel (CSR by mel bin)
    int n_mel = 0, mel_f_lo = 0, mel_f_n = 0, mel_taps = 0;
    int mel_L[4] = {0, 0, 0, 0};
    DevBuf mel_info, mel_w;
    // plan
    bool has_plan = false, labels_done = false;
    int B = 0, T = 0, V = 0, M = 0, C = 0;
    int n_tmask = 0, n_fmask = 0, filter_k = 0, remap = 0, c_out = 0;
    std::vector<Seg> h_segs;
    std::vector<int64_t> h_seg_len;   // true samples per channel of each segment's source
    std::vector<int32_t> h_seg_ptr;
    DevBuf keep, minmax, scratch_labels, stft_pad, stft_small;   // minmax: [B,2] + done [B]
    DevBuf tiles, sched;
    int max_segs = 1;
    // tile blocks built by k_labels for the feature launch that follows (iris_abi.cu: run_labels):
    // the layout they were built for; a feature launch with another one runs k_tiles
    struct TileSig {
        bool valid = false;
        const void *segs = nullptr, *keep = nullptr;   // the plan (a plan upload invalidates as well)
        int B = 0, T = 0;
        int fm_bits = 0, seg_select = 0, fr = 0, n_pairs = 0, max_segs = 0, stride = 0, masks = 0, filter_k = 0;
    } tile_sig;
    int feat_hint = -1;   // feature mode of the last feature launch: what k_labels builds tile blocks for
    cudaEvent_t ev_after_fused = nullptr;   // recorded (once) right behind the next k_fused launch (iris_step)
    // pinned staging for the plan blob: a ring, so that the host can run up to kStageRing batches
    // ahead of the device before it has to wait for an upload to drain
    static constexpr int kStageRing = 4;
    void* h_stage[kStageRing] = {nullptr, nullptr, nullptr, nullptr};
    size_t h_stage_cap[kStageRing] = {0, 0, 0, 0};
    cudaEvent_t stage_free[kStageRing] = {nullptr, nullptr, nullptr, nullptr};
    int stage_next = 0;
    size_t last_upload_bytes = 0;
    DevBuf plan_blobs[kStageRing];     // device side of the ring
    cudaStream_t copy = nullptr;       // plan uploads
    cudaEvent_t blob_ready[kStageRing] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t blob_used[kStageRing] = {nullptr, nullptr, nullptr, nullptr};
    bool blob_used_valid[kStageRing] = {false, false, false, false};
    // device views into plan_blob
    Seg* d_segs = nullptr;
    int32_t *d_seg_ptr = nullptr, *d_n_voices = nullptr, *d_voice_id = nullptr,
            *d_voice_shift = nullptr, *d_voice_kt = nullptr, *d_tmask = nullptr, *d_fmask = nullptr;
    float *d_merge_f = nullptr, *d_merge_sf = nullptr;
    // roofline measurement hook
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    int prof_clips = 0;               // clips of the launch the hook timed last (a split batch: its first part)
    // stand-alone ops (iris_ops_abi.cu): dense mel matrix + column supports, small scratch
    DevBuf mel_dense, mel_lo, mel_len, op_small, minmax_ops, eval_scratch, spec_scratch;
    bool spec_mode = false;            // the uploaded plan mixes spectrogram banks
    int mel_bins = 0;
    bool mel_fusable = false;
    // split feature launches (iris_abi.cu): parts of a large batch alternate between the caller's
    // stream and `aux`, so that k_logmel_post of one part runs beside k_fused of the next
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_tiles = nullptr, ev_join = nullptr;
    // ---- one-call step (iris_step.cu) ----
    cudaStream_t side = nullptr;       // metric leg: counting + count all-reduce beside the feature kernel
    cudaEvent_t ev_labels = nullptr;   // labels of the current step are written
    static constexpr int kLegRing = 8;
    struct MetricLeg {
        cudaEvent_t done = nullptr;
        const void* labels = nullptr;  // frame-label buffer the leg reads
        uint64_t seq = 0;              // iris_step call that issued it (1-based); 0: never used
    } legs[kLegRing];
    uint64_t step_seq = 0;
    std::vector<int32_t> draw_i32;     // host draws of the last iris_step (views in `draws`)
    std::vector<float> draw_f32;
    iris_draws draws{};
    int mel_read_wavefronts = 0;       // modelled shared-memory wavefronts per frame of the mel tap reads
};

int iris_set_device(iris_ctx* c);
void iris_step_release(iris_ctx* c);   // iris_step.cu: side stream + events
