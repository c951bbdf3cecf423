cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2e
for cs in "32 128 9600 11 5" "32 128 9600 3 1" "32 256 2400 7 1" "32 64 19200 11 1" "32 64 19200 3 1" "32 32 38400 7 1"; do
  TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1
done > gpurun_out/r2e/harness.txt 2>&1
cat gpurun_out/r2e/harness.txt | cut -c1-60,100-260
python -m pytest tests/test_ragged_gpu.py tests/test_handoff_gpu.py tests/test_mas_gpu.py -m gpu -q > gpurun_out/r2e/pytest_new.txt 2>&1; tail -25 gpurun_out/r2e/pytest_new.txt
python -m pytest tests -m gpu -q --deselect tests/test_ragged_gpu.py --deselect tests/test_handoff_gpu.py --deselect tests/test_mas_gpu.py > gpurun_out/r2e/pytest_rest.txt 2>&1; tail -6 gpurun_out/r2e/pytest_rest.txt
python scripts/dev_bench_mas.py > gpurun_out/r2e/mas.txt 2>&1; tail -2 gpurun_out/r2e/mas.txt
B200TTS_MAS2=1 python scripts/dev_bench_mas.py 2>&1 | tail -2
BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ragged', d['ms_per_step'], d['value'], d['config']['stage_ms_per_step'], d['roofline']['achieved'])"
