cd $GRAFT_REPO_ROOT
for cs in "2 128 700 11 5" "2 128 1004 7 3" "1 256 300 3 1" "3 192 600 5 1" "2 512 520 7 1" "2 64 1004 11 5"; do
  for ra in "1 0" "0 1"; do set -- $ra
    TC_V3=1 TC_G=1 timeout 60 ./tools/test_conv_tc one $cs 0 $1 $2 2>&1 | tail -1 | cut -c1-60,95-175
  done
done
for cs in "32 128 9600 11 5" "32 256 1100 7 1" "32 192 192 5 1"; do
  echo -n "$cs: "; TC_V3=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms" | tr '\n' ' '
  echo -n " full-N: "; TC_DBG=512 TC_V3=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['stage_ms_per_step'])"
