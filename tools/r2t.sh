# activation-stage depth variants (NA2 = transformed A stages, NB2 = weight ring slots)
cd $GRAFT_REPO_ROOT
for v in "d NA3_NB10" "g NA4_NB8" "h NA5_NB6" "i NA4_NB10"; do set -- $v
  echo "== $2"
  for cs in "32 128 9600 11 5" "32 128 9600 7 3" "32 128 9600 3 1" "32 256 2400 3 1" "32 64 19200 3 1" "32 64 19200 11 1" "32 32 38400 3 1" "32 32 38400 7 1"; do
    echo -n "  $cs : "; TC_V3=1 TC_G=1 ./tools/test_conv_tc_$1 one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms\|KERNEL FAILED.*\|invalid.*" | tr '\n' ' '; echo
  done
done
