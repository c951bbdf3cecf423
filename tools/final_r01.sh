# Round-end verification + profile capture (run under gpurun from the repo root)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/final
echo "### pytest"; python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "### smoke"; python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "### bench"; python bench.py > gpurun_out/final/bench_1gpu.json 2> gpurun_out/final/bench_1gpu.err; cat gpurun_out/final/bench_1gpu.json | head -c 600; echo
echo "### launch list"; ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/final/bench_launches.csv env BENCH_SKIP_CPU=1 python bench.py --steps 1 --warmup 3 > /dev/null 2>&1; wc -l gpurun_out/final/bench_launches.csv
echo "### decoder dram traffic"; ITERS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final/decoder_dram.csv python scripts/dev_bench_hifigan.py 32 192 > gpurun_out/final/decoder_dram.out 2>&1; wc -l gpurun_out/final/decoder_dram.csv
echo "### ncu full, grouped kernel"; ncu --set full --import-source on --clock-control none -k regex:conv1d_tc3g -c 1 -o /tmp/g4f -f env TC_V3=1 TC_G=1 ./tools/test_conv_tc one 32 32 38400 7 1 0 > /dev/null 2>&1
ncu -i /tmp/g4f.ncu-rep --page details --csv > gpurun_out/final/ncu_tc3g_c32k7.details.csv 2>&1
ncu -i /tmp/g4f.ncu-rep --page raw --csv > gpurun_out/final/ncu_tc3g_c32k7.raw.csv 2>&1
echo "### sanitizer"; (
for cs in "2 32 724 3 5" "2 64 1004 11 5" "2 32 2000 7 1" "2 64 724 3 3"; do TC_V3=1 TC_G=1 compute-sanitizer --tool memcheck ./tools/test_conv_tc one $cs 0 2>&1 | grep -E "ERROR SUMMARY|OK|MISMATCH" | tr '\n' ' '; echo " [grouped $cs]"; done
TC_V3=1 compute-sanitizer --tool memcheck ./tools/test_conv_tc one 2 128 700 11 5 0 2>&1 | grep -E "ERROR SUMMARY|OK|MISMATCH" | tr '\n' ' '; echo " [plain v3]"
compute-sanitizer --tool memcheck python -m pytest tests/test_vits_layers_gpu.py tests/test_hifigan_gpu.py -x -q -k "golden" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tr '\n' ' '; echo " [pytest golden]"
compute-sanitizer --tool memcheck python -m pytest tests/test_vits_e2e_gpu.py -x -q -k "voice_conversion or deterministic or single_speaker" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tr '\n' ' '; echo " [pytest e2e]"
) > gpurun_out/final/sanitizer.txt 2>&1; cat gpurun_out/final/sanitizer.txt
