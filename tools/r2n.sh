cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2n
for cs in "32 128 9600 3 1" "32 32 38400 3 1" "32 64 19200 3 1"; do
  tag=$(echo $cs | tr ' ' '_')
  ncu --set full --import-source on --clock-control none -k regex:conv1d_tc3 -c 1 -o gpurun_out/r2n/full_$tag -f env TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 0 > /dev/null 2>&1
  echo "-- trace $cs"; TC_TRACE=1 TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 0 2>&1 | grep -A3 "v3 trace" | tail -2
done
BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['stage_ms_per_step'])"
ls -la gpurun_out/r2n
