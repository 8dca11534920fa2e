mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest12.log
cat gpurun_out/pytest12.log
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); st=d['stages']
        print('value %.3e e2e %.3e ms/step %.3f | pose %.1f gemm %.1f skin %.1f score %.1f us/step  clk %s %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], st['pose_chain']['ms_total']/d['steps']*1e3, st['blend_gemm']['ms_total']/d['steps']*1e3, st['skinning']['ms_total']/d['steps']*1e3, st['scoring']['ms_total']/d['steps']*1e3, d['clocks']['sm_mhz'], d['clocks']['reasons']))
    elif 'rror' in l or 'Trace' in l: print(l.strip())
"; }
PRK_BENCH_PRELOAD_S=0.3 timeout 120 python bench.py --steps 50 --warmup 5 2>&1 | show | tee gpurun_out/sweep12.log
