# pipeline-depth variants of the harness: NRAW (cp.async ring) / NB2 (weight ring) / NPW (producer warps) / RDEPTH (residual sets)
cd $GRAFT_REPO_ROOT
for v in "d NRAW4_NB10_NPW6_RD2" "c NRAW4_NB10_NPW4_RD2" "a NRAW6_NB6_NPW6_RD2" "b NRAW8_NB4_NPW6_RD2" "e NRAW6_NB6_NPW6_RD4" "f NRAW8_NB4_NPW6_RD4"; do set -- $v
  echo "== $2"
  for cs in "32 128 9600 11 5" "32 128 9600 7 3" "32 128 9600 3 1" "32 256 2400 3 1" "32 64 19200 3 1" "32 64 19200 11 1" "32 32 38400 3 1" "32 32 38400 7 1"; do
    echo -n "  $cs : "; TC_V3=1 TC_G=1 ./tools/test_conv_tc_$1 one $cs 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms\|KERNEL FAILED.*" | tr '\n' ' '; echo
  done
  for dbg in 48 49 50 12; do echo -n "  32 128 9600 3 1 dbg=$dbg : "; TC_DBG=$dbg TC_V3=1 TC_G=1 ./tools/test_conv_tc_$1 one 32 128 9600 3 1 10 2>&1 | tail -1 | grep -o "OK *[0-9.]* ms\|MISMATCH *[0-9.]* ms"| tr '\n' ' '; echo; done
done
