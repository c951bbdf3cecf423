# narrow-layer role ablation (TC_DBG: 1 no cp.async, 2 no transform, 4 no epilogue loads, 8 no stores, 16 no MMA) + racecheck details
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2k
for cs in "32 32 38400 3 1" "32 32 38400 11 1" "32 64 19200 7 1" "32 128 9600 3 1"; do
  for dbg in 0 1 2 3 4 8 12 16 19 28 31; do
    echo -n "$cs dbg=$dbg : "; TC_DBG=$dbg TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms"| tr '\n' ' '; echo
  done
done
for cs in "2 128 700 11 5" "3 64 1004 7 3" "2 32 2000 3 1"; do TC_V3=1 TC_G=1 compute-sanitizer --tool racecheck --print-limit 30 ./tools/test_conv_tc one $cs 0 > gpurun_out/r2k/race_harness_$(echo $cs | tr ' ' _).txt 2>&1; grep -E "RACECHECK SUMMARY" gpurun_out/r2k/race_harness_*.txt | tail -1; done
compute-sanitizer --tool racecheck --print-limit 40 python -m pytest tests/test_ragged_gpu.py -x -q -k "flow or peak" > gpurun_out/r2k/race_pytest.txt 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2k/race_pytest.txt
grep -E "hazard detected|^=========     at " gpurun_out/r2k/race_pytest.txt | sort | uniq -c | sort -rn | head -20
