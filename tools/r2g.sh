cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2g
for cs in "32 128 9600 3 1" "32 128 9600 11 5" "32 64 19200 3 1" "32 32 38400 7 1" "32 32 38400 3 1"; do
  echo "== $cs"; TC_TRACE=1 TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 3 2>&1 | tail -4
done > gpurun_out/r2g/trace.txt 2>&1
cat gpurun_out/r2g/trace.txt | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2g/bench_launches.csv env BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py --steps 1 --warmup 3 > /dev/null 2>&1; wc -l gpurun_out/r2g/bench_launches.csv
