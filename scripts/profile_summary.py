"""Turn the ncu outputs of scripts/gpu_profile.sh (gpurun_out/) into the tracked summaries under profiles/:
python scripts/profile_summary.py <tag>   ->  profiles/<tag>_launches.csv, <tag>_launch_summary.txt, <tag>_fused_raw.csv,
<tag>_fused_stalls.txt and profiles/fused_traffic.json."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, 'profiles')
src = os.path.join(ROOT, 'gpurun_out')

# 1. launch list
shutil.copy(os.path.join(src, 'launches.csv'), os.path.join(out, f'{tag}_launches.csv'))
rows = [r for r in csv.reader(open(os.path.join(src, 'launches.csv'))) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split('(')[0].replace('prk::', '').replace('<unnamed>::', '')
    tot[name][0] += 1
    tot[name][1] += float(r[-1].replace(',', '')) / 1e3
total = sum(v[1] for v in tot.values())
with open(os.path.join(out, f'{tag}_launch_summary.txt'), 'w') as f:
    f.write('ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 3 --warmup 3 --repeats 1 --skip-extra (PRK_BENCH_PRELOAD_S=0)\n')
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write('%-60s launches %3d  total %8.1f us  share %5.1f%%  avg %6.1f us\n' % (name, n, us, 100 * us / total, us / n))
print(open(os.path.join(out, f'{tag}_launch_summary.txt')).read())

# 2. full capture of the fused kernel: raw page, stall reasons, DRAM traffic
rep = os.path.join(src, 'prof_fused.ncu-rep')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
open(os.path.join(out, f'{tag}_fused_raw.csv'), 'w').write(raw)
r = list(csv.reader(raw.splitlines()))
hdr, unit, val = r[0], r[1], r[2]
d = {h: v for h, v in zip(hdr, val)}
u = {h: x for h, x in zip(hdr, unit)}


def num(k):
    return float(d[k].replace(',', ''))


def to_bytes(k):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u[k]]
    return num(k) * scale


rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
json.dump({"kernel": "fused_blend_skin_kernel", "source": f"profiles/{tag}_fused_raw.csv (ncu --set full, 4096 frames per launch)",
           "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr},
          open(os.path.join(out, 'fused_traffic.json'), 'w'), indent=1)
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
with open(os.path.join(out, f'{tag}_fused_stalls.txt'), 'w') as f:
    f.write('fused_blend_skin_kernel<1>, one launch of 4096 frames (ncu --set full --clock-control none)\n')
    for k in keys:
        if k in d:
            f.write('%-86s %s %s\n' % (k, d[k], u[k]))
    f.write('dram read %.1f MB, write %.1f MB per launch\n' % (rd / 1e6, wr / 1e6))
    st = [(num(h), h) for h in hdr if 'smsp__pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and d[h]]
    s = sum(x for x, _ in st)
    f.write('warp stall samples (all warps):\n')
    for x, h in sorted(st, reverse=True)[:10]:
        f.write('  %5.1f%%  %s\n' % (100 * x / s, h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
print(open(os.path.join(out, f'{tag}_fused_stalls.txt')).read())
