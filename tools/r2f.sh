cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2f
for cs in "32 128 9600 11 5" "32 128 9600 3 1" "32 256 2400 7 1" "32 64 19200 11 1" "32 64 19200 3 1" "32 32 38400 7 1" "2 128 700 11 5" "3 64 1004 7 3"; do
  TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1
done > gpurun_out/r2f/harness.txt 2>&1
cat gpurun_out/r2f/harness.txt | cut -c1-60,100-260
python -m pytest tests -m gpu -q -x > gpurun_out/r2f/pytest.txt 2>&1; tail -5 gpurun_out/r2f/pytest.txt
BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ragged', d['ms_per_step'], d['value'], d['config']['stage_ms_per_step'], d['roofline']['achieved'])"
