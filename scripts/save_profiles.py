#!/usr/bin/env python
"""Turn the artefacts of scripts/gpu_round.sh (gpurun_out/) into the tracked summaries under
profiles/.  usage: save_profiles.py TAG "description of the kernel revision" """
import collections, csv, os, shutil, subprocess, sys
tag, desc = sys.argv[1], sys.argv[2]
G, P = 'gpurun_out', 'profiles'
rep = os.path.join(G, 'prof_fused.ncu-rep')
def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout
# launch list
rows = [r for r in csv.reader(open(os.path.join(G, 'launches.csv'))) if len(r) > 10]
h = rows[0]; agg = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(h, r))
    if d['Metric Name'] != 'gpu__time_duration.sum': continue
    agg.setdefault(d['Kernel Name'], []).append(float(d['Metric Value'].replace(',', '')) / 1e3)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(P, '%s_launches_summary.txt' % tag), 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e (%s)\n' % desc)
    f.write('# includes bank registration (k_bank_prepare, k_fused<4> activity pass) and torch fills; per-launch times are cold-cache and serialised\n')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write('%-70s n=%3d total_us=%10.1f avg_us=%9.1f share=%.3f\n' % (k[:70], len(v), sum(v), sum(v) / len(v), sum(v) / tot))
# details / raw
det = run(['ncu', '-i', rep, '--page', 'details'])
open(os.path.join(P, '%s_fused_ncu_details.txt' % tag), 'w').write('# ncu --set full --clock-control none --import-source on -k regex:k_fused -s 4 -c 2: python bench.py --steps 2 --warmup 3 (%s)\n' % desc + det)
raw = list(csv.reader(run(['ncu', '-i', rep, '--page', 'raw', '--csv']).splitlines()))
with open(os.path.join(P, '%s_fused_ncu_raw.txt' % tag), 'w') as f:
    f.write('# ncu --set full --clock-control none -k regex:k_fused: python bench.py --steps 2 --warmup 3 (%s); first captured launch\n' % desc)
    for name, unit, val in zip(raw[0], raw[1], raw[2]):
        f.write('%-90s %s %s\n' % (name, val, unit))
src = os.path.join('/tmp', 'src_page_%s.csv' % tag)
open(src, 'w').write(run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass']))
here = os.path.dirname(os.path.abspath(__file__))
open(os.path.join(P, '%s_fused_opcodes.txt' % tag), 'w').write('# ncu source page, k_fused<FM_MEL,4> (%s): executed instructions by opcode\n' % desc + run([sys.executable, os.path.join(here, 'ncu_ops.py'), src]))
open(os.path.join(P, '%s_fused_source_stalls.txt' % tag), 'w').write('# ncu source page, k_fused<FM_MEL,4> (%s): stall samples / instructions by source line\n' % desc + run([sys.executable, os.path.join(here, 'ncu_lines.py'), src, '40']))
shutil.copy(os.path.join(G, 'bench.json'), os.path.join(P, '%s_bench.json' % tag))
if os.path.exists(os.path.join(G, 'bench_ref.json')):
    shutil.copy(os.path.join(G, 'bench_ref.json'), os.path.join(P, '%s_bench_reference.json' % tag))
if os.path.exists(os.path.join(G, 'kbench.log')):
    shutil.copy(os.path.join(G, 'kbench.log'), os.path.join(P, '%s_kbench_modes.txt' % tag))
print('saved', tag)
