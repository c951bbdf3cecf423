# racecheck probe: the same pipeline with (a) tcgen05.commit arrivals, MMAs off and (b) plain thread arrivals, MMAs off
cd $GRAFT_REPO_ROOT
# build the harness with -DTC3_RACE_PROBE for the 272 case
for dbg in 0 16 272; do TC_DBG=$dbg TC_V3=1 compute-sanitizer --tool racecheck --print-limit 2 ./tools/test_conv_tc one 2 128 700 11 5 0 2>&1 | grep -E "RACECHECK SUMMARY|hazards\]" | cut -c1-160 | tr '\n' ' '; echo " [dbg=$dbg]"; done
