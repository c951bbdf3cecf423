# lean direct epilogue + quadrant grouped epilogue + lean producers (6 warps): correctness + timing + parity tests + bench
cd $GRAFT_REPO_ROOT
echo "## correctness, small / edge shapes (res, accum variants)"
for cs in "2 32 724 3 5" "2 64 1004 11 5" "2 32 2000 7 1" "3 64 1004 7 3" "2 128 700 11 5" "1 256 300 3 1" "2 128 1004 7 3" "3 192 600 5 1" "2 512 520 7 1"; do
  for ra in "1 0" "0 0" "1 1"; do set -- $ra
    TC_V3=1 TC_G=1 timeout 60 ./tools/test_conv_tc one $cs 0 $1 $2 2>&1 | tail -1 | cut -c1-60,95-175
  done
done
echo "## timing"
for cs in "32 128 9600 11 5" "32 128 9600 7 3" "32 128 9600 3 1" "32 256 2400 7 1" "32 256 2400 3 1" "32 64 19200 11 1" "32 64 19200 3 1" "32 32 38400 7 1" "32 32 38400 3 1"; do
  TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1 | cut -c1-50,110-220
done
for cs in "32 32 38400 3 1" "32 128 9600 3 1"; do for dbg in 12 31 32 48; do echo -n "$cs dbg=$dbg : "; TC_DBG=$dbg TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms"| tr '\n' ' '; echo; done; done
for cs in "32 128 9600 3 1" "32 64 19200 3 1"; do echo "-- trace $cs"; TC_TRACE=1 TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 0 2>&1 | grep -A3 "v3 trace" | tail -2; done
echo "## pytest"
python -m pytest tests/test_hifigan_gpu.py tests/test_ragged_gpu.py tests/test_bench_scale_gpu.py tests/test_vits_layers_gpu.py -x -q 2>&1 | tail -4
BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['stage_ms_per_step'])"
