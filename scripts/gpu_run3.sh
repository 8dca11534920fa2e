mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/pytest3.log
cat gpurun_out/pytest3.log
timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/bench3.log 2>&1; tail -1 gpurun_out/bench3.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); st=d['stages']
        print('value %.3e e2e %.3e ms/step %.3f | pose %.1f gemm %.1f skin %.1f score %.1f us/step' % (d['value'], d['e2e']['value'], d['ms_per_step'], st['pose_chain']['ms_total']/d['steps']*1e3, st['blend_gemm']['ms_total']/d['steps']*1e3, st['skinning']['ms_total']/d['steps']*1e3, st['scoring']['ms_total']/d['steps']*1e3)); print(d['roofline']); print(d['roofline_other']); print(d['clocks'])
    else: print(l)
"
