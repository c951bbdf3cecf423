cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for v in test_conv_tc_old test_conv_tc; do
  for cs in "32 128 9600 11 5" "32 256 1200 7 3" "32 128 9600 3 1"; do
    echo "== $v plain case $cs"; TC_V3=1 timeout 60 ./tools/$v one $cs 10 2>&1 | grep -o "MISMATCH.*\|OK .*"
  done
done; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv
