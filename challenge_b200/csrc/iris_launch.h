// Host-callable launchers of the libiris kernels (internal; the public surface is include/iris.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iris {

struct FusedParams;

cudaError_t launch_fused(const FusedParams& p, int mode, int num_sms, cudaStream_t stream);
size_t fused_smem_bytes(const FusedParams& p, int mode);
int fused_max_segments();      // mixing segments one clip may have
int fused_max_mel_window();    // widest bin range [f_lo, f_hi] of the mel matrix
int fused_max_mel_taps();      // sum over 16-filter groups of the longest filter
size_t fused_tile_bytes(const FusedParams& p, int* stride_out);   // size of p.tile_blocks

// k_post.cu
cudaError_t launch_logmel_post(float* x, uint32_t* minmax, int B, size_t per_clip, int do_minmax,
                               int do_log, cudaStream_t stream);

// k_bank.cu
cudaError_t launch_bank_prepare(const float* wav, const int64_t* d_offsets, const int64_t* d_pad_offsets,
                                int n_items, int n_chan, int normalize, float* padded,
                                cudaStream_t stream);

// k_labels.cu
struct LabelParams {
    int B, T, V, K;
    const int32_t* n_voices;     // [B]
    const int32_t* voice_id;     // [B,V]
    const int32_t* voice_shift;  // [B,V]  source frame k = t + shift
    const int32_t* n_frames;     // [n_items] true frame count of each voice
    const uint8_t* activity;     // [n_items, act_stride]
    int act_stride;
    const float* bank_labels;    // [n_items, K]
    float* labels_vtk;           // [B,V,T,K] or null
    float* frame_labels;         // [B,T,K]
    uint8_t* keep;               // [B,V]
};
cudaError_t launch_labels(const LabelParams& p, cudaStream_t stream);

// k_metrics.cu
cudaError_t launch_metric_counts(const float* y_true, const float* y_pred, int B, int T, int K,
                                 float threshold, int32_t* triples, unsigned long long* tpfpfn,
                                 unsigned long long* sums, cudaStream_t stream);
cudaError_t launch_er_finalize(const int32_t* triples, int B, float* er, cudaStream_t stream);

}  // namespace iris
