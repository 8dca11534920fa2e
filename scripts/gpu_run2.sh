mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "blend or smpl or pipeline or host_path" 2>&1 | tail -5 > gpurun_out/pytest2.log
cat gpurun_out/pytest2.log
for c in 384 640 896 1024 1280 2048 4096; do
  echo "chunk $c"
  PRK_CHUNK_FRAMES=$c timeout 120 python bench.py --steps 30 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); st=d['stages']
        print('value %.3e e2e %.3e ms/step %.3f | pose %.1f gemm %.1f skin %.1f score %.1f us/step | gemm exec frac %.3f skin frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], st['pose_chain']['ms_total']/d['steps']*1e3, st['blend_gemm']['ms_total']/d['steps']*1e3, st['skinning']['ms_total']/d['steps']*1e3, st['scoring']['ms_total']/d['steps']*1e3, (d['roofline_other'] if d['roofline']['bound']=='hbm' else d['roofline'])['executed_mma']['frac'], (d['roofline'] if d['roofline']['bound']=='hbm' else d['roofline_other'])['frac']))
    elif 'rror' in l: print(l.strip())
"
done 2>&1 | tee gpurun_out/sweep2.log
