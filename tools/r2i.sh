cd $GRAFT_REPO_ROOT
for cs in "32 128 9600 11 5" "32 128 9600 7 3" "32 256 2400 7 1" "32 256 2400 11 5"; do
  for dbg in 0 32 16 48; do
    echo -n "$cs dbg=$dbg : "; TC_DBG=$dbg TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH.*"
  done
done
