cd $GRAFT_REPO_ROOT
for bin in test_conv_tc_nb6 test_conv_tc test_conv_tc_nb14; do
for cs in "32 128 9600 11 5" "32 128 9600 7 3" "32 128 9600 3 1" "32 256 2400 7 1" "32 64 19200 11 1" "32 32 38400 3 1"; do
  for dbg in 0 16; do
    echo -n "$bin $cs dbg=$dbg : "; TC_DBG=$dbg TC_V3=1 TC_G=1 ./tools/$bin one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms"| tr '\n' ' '; echo
  done
done; done
for cs in "2 128 700 11 5" "3 64 1004 7 3" "2 32 2000 3 1" "1 256 300 3 1"; do TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 2 2>&1 | tail -1 | cut -c1-40,100-200; done
