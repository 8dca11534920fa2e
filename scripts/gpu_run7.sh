mkdir -p gpurun_out
export PRK_BENCH_PRELOAD_S=0
CMD="python bench.py --steps 3 --warmup 3"
$CMD > gpurun_out/plain7.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu7.log 2>&1
$CMD > gpurun_out/plain7b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:skin_mma_kernel -s 3 -c 1 -o gpurun_out/prof_skinmma_r1b $CMD > gpurun_out/ncu7b.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_r1b.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:40]:
    print(r[ki][:70], r[vi])
PY
