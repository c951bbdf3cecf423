# Round-2 verification + profile capture (run under gpurun from the repo root)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/final2
echo "### pytest"; python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "### smoke"; python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "### bench"; python bench.py > gpurun_out/final2/bench_1gpu.json 2> gpurun_out/final2/bench_1gpu.err; head -c 900 gpurun_out/final2/bench_1gpu.json; echo
echo "### bench dense"; BENCH_DENSE=1 BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py > gpurun_out/final2/bench_1gpu_dense.json 2>/dev/null; head -c 400 gpurun_out/final2/bench_1gpu_dense.json; echo
echo "### launch list"; ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/final2/bench_launches.csv env BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py --steps 1 --warmup 3 > /dev/null 2>&1; wc -l gpurun_out/final2/bench_launches.csv
echo "### decoder dram traffic (dense B=32 x 192 frames, 1 pass after 2 warm-ups)"; ITERS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final2/decoder_dram.csv python scripts/dev_bench_hifigan.py 32 192 > gpurun_out/final2/decoder_dram.out 2>&1; wc -l gpurun_out/final2/decoder_dram.csv
echo "### ncu full, top kernel (C=128 K=11, the stage-1 MRF layer)"; ncu --set full --import-source on --clock-control none -k regex:conv1d_tc3_kernel -c 1 -o gpurun_out/final2/tc3_c128k11 -f env TC_V3=1 ./tools/test_conv_tc one 32 128 9600 11 5 0 > /dev/null 2>&1
ncu -i gpurun_out/final2/tc3_c128k11.ncu-rep --page details --csv > gpurun_out/final2/ncu_tc3_c128k11.details.csv 2>&1
ncu -i gpurun_out/final2/tc3_c128k11.ncu-rep --page raw --csv > gpurun_out/final2/ncu_tc3_c128k11.raw.csv 2>&1
echo "### ncu full, MAS"; ncu --set full --clock-control none -k regex:mas_kernel2 -c 1 -o gpurun_out/final2/mas2 -f env ITERS=1 python scripts/dev_bench_mas.py > /dev/null 2>&1
ncu -i gpurun_out/final2/mas2.ncu-rep --page details --csv > gpurun_out/final2/ncu_mas2.details.csv 2>&1
echo "### stft timing"; python scripts/dev_bench_stft.py 2>&1 | tail -3
echo "### sanitizer"; (
for cs in "2 128 700 11 5" "3 64 1004 7 3" "2 32 2000 3 1"; do TC_V3=1 TC_G=1 compute-sanitizer --tool memcheck ./tools/test_conv_tc one $cs 0 2>&1 | grep -E "ERROR SUMMARY|OK|MISMATCH" | tr '\n' ' '; echo " [memcheck harness $cs]"; done
for cs in "2 128 700 11 5" "3 64 1004 7 3" "2 32 2000 3 1"; do TC_V3=1 TC_G=1 compute-sanitizer --tool synccheck ./tools/test_conv_tc one $cs 0 2>&1 | grep -E "ERROR SUMMARY|OK|MISMATCH" | tr '\n' ' '; echo " [synccheck harness $cs]"; done
compute-sanitizer --tool memcheck python -m pytest tests/test_ragged_gpu.py tests/test_handoff_gpu.py -x -q 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tr '\n' ' '; echo " [memcheck pytest ragged+handoff]"
compute-sanitizer --tool racecheck python -m pytest tests/test_handoff_gpu.py tests/test_mas_gpu.py -x -q 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed" | tr '\n' ' '; echo " [racecheck pytest handoff+mas: kernels without tcgen05]"
for cs in "3 64 1004 7 3" "2 32 2000 3 1" "2 128 700 11 5"; do TC_V3=1 TC_G=1 compute-sanitizer --tool racecheck --print-limit 3 ./tools/test_conv_tc one $cs 0 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|and .* access|OK|MISMATCH" | cut -c1-150 | tr '\n' ' '; echo " [racecheck harness $cs]"; done
) > gpurun_out/final2/sanitizer.txt 2>&1; cat gpurun_out/final2/sanitizer.txt
