cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2h
for cs in "32 128 9600 11 5" "32 128 9600 3 1" "32 128 9600 7 3" "32 64 19200 3 1" "32 32 38400 7 1"; do
  for dbg in 0 4 8 12 3 15 16 28; do
    echo -n "$cs dbg=$dbg : "; TC_DBG=$dbg TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 10 2>&1 | tail -1 | grep -o "[0-9.]* ms" 
  done
done > gpurun_out/r2h/dbg.txt 2>&1
cat gpurun_out/r2h/dbg.txt
