set -x
cd $GRAFT_REPO_ROOT
for dbg in 0 1 2 3 4 8 12 16 31; do
  for cs in "32 32 38400 3 1" "32 64 19200 3 1" "32 32 38400 7 1"; do
    echo "== dbg=$dbg case $cs"
    TC_V3=1 TC_G=1 TC_DBG=$dbg timeout 60 ./tools/test_conv_tc one $cs 5 2>&1 | grep -o "MISMATCH.*\|OK .*"
  done
done
TC_V3=1 TC_DBG=0 timeout 60 ./tools/test_conv_tc one 32 128 9600 3 1 5 | grep -o "OK .*"
TC_V3=1 TC_DBG=1 timeout 60 ./tools/test_conv_tc one 32 128 9600 3 1 5 | grep -o "MISMATCH.*\|OK .*"
TC_V3=1 TC_DBG=2 timeout 60 ./tools/test_conv_tc one 32 128 9600 3 1 5 | grep -o "MISMATCH.*\|OK .*"
TC_V3=1 TC_DBG=16 timeout 60 ./tools/test_conv_tc one 32 128 9600 3 1 5 | grep -o "MISMATCH.*\|OK .*"
TC_V3=1 TC_G=1 TC_TRACE=1 timeout 60 ./tools/test_conv_tc one 32 32 38400 3 1 0
