# skeleton probes: where do the ~1500 cycles per chunk go?  traces + epilogue-handshake-only probe + ncu source of two K=3 layers
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2l
for cs in "32 32 38400 3 1" "32 128 9600 3 1"; do
  for dbg in 0 31 32 48 63; do
    echo -n "$cs dbg=$dbg : "; TC_DBG=$dbg TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms"| tr '\n' ' '; echo
  done
  for dbg in 0 31 63; do echo "-- trace dbg=$dbg"; TC_TRACE=1 TC_DBG=$dbg TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 0 2>&1 | grep -A3 "v3 trace" | tail -3; done
  tag=$(echo $cs | tr ' ' '_')
  ncu --set full --import-source on --clock-control none -k regex:conv1d_tc3 -c 1 -o gpurun_out/r2l/full_$tag -f env TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 0 > /dev/null 2>&1
done
echo "-- racecheck with the test_wait probe disabled (dbg 64)"
TC_DBG=64 TC_V3=1 compute-sanitizer --tool racecheck --print-limit 5 ./tools/test_conv_tc one 2 128 700 11 5 0 2>&1 | grep -E "RACECHECK SUMMARY|OK|MISMATCH|hazards\]" | head -5
TC_V3=1 compute-sanitizer --tool racecheck --print-limit 5 ./tools/test_conv_tc one 2 128 700 11 5 0 2>&1 | grep -E "RACECHECK SUMMARY|OK|MISMATCH|hazards\]" | head -5
ls -la gpurun_out/r2l
