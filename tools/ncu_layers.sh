# ncu --set full on two representative layers through the stand-alone harness (run under gpurun from the repo root):
#   C=128 K=11 d=5 (stage-1 MRF, MMA-heavy) and C=64 K=3 d=1 grouped (epilogue / traffic heavy)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/ncu
M="gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__cycles_active.avg,sm__cycles_elapsed.avg,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_bytes.sum.per_second,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active"
for cs in "32 128 9600 11 5" "32 128 9600 3 1" "32 64 19200 3 1" "32 32 38400 7 1"; do
  tag=$(echo $cs | tr ' ' '_')
  ncu --metrics $M --clock-control none -k regex:conv1d_tc3 -c 1 --csv --log-file gpurun_out/ncu/m_$tag.csv env TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 0 > /dev/null 2>&1
  ncu --set full --import-source on --clock-control none -k regex:conv1d_tc3 -c 1 -o gpurun_out/ncu/full_$tag -f env TC_V3=1 TC_G=1 ./tools/test_conv_tc one $cs 0 > /dev/null 2>&1
  ncu -i gpurun_out/ncu/full_$tag.ncu-rep --page details --csv > gpurun_out/ncu/details_$tag.csv 2>&1
done
ls -la gpurun_out/ncu
