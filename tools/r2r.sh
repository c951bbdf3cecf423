cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2r
BENCH_DENSE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2r/dense_launches.csv env BENCH_DENSE=1 BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2r/ragged_launches.csv env BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py --steps 1 --warmup 3 > /dev/null 2>&1
wc -l gpurun_out/r2r/*.csv
