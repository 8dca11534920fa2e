for f in 0 1 64 255 245; do
  PRK_SKIN_DBGFLAGS=$f PRK_BENCH_PRELOAD_S=0.3 timeout 100 python bench.py --steps 20 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); st=d['stages']; print('flags $f: skin %.1f us/step  gemm %.1f' % (st['skinning']['ms_total']/d['steps']*1e3, st['blend_gemm']['ms_total']/d['steps']*1e3))
    elif 'rror' in l: print(l.strip()[:200])
"
done
