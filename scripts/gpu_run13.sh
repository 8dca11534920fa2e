mkdir -p gpurun_out
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); st=d['stages']
        print('ms/step %.3f | pose %.1f fused %.1f score %.1f us/step  clk %s %s' % (d['ms_per_step'], st['pose_chain']['ms_total']/d['steps']*1e3, st['blend_gemm']['ms_total']/d['steps']*1e3, st['scoring']['ms_total']/d['steps']*1e3, d['clocks']['sm_mhz'], d['clocks']['reasons']))
    elif 'rror' in l or 'Trace' in l: print(l.strip())
"; }
for f in ${FLAGS:-0 1 2 3}; do
  echo "dbg $f"
  PRK_FUSED_DBG=$f PRK_BENCH_PRELOAD_S=0.2 timeout 120 python bench.py --steps 30 --warmup 3 2>&1 | show
done 2>&1 | tee gpurun_out/sweep13.log
