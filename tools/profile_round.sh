#!/bin/bash
# ncu evidence of one round (run on the GPU box through gpurun):  bash tools/profile_round.sh r02
#   <tag>_ncu_launches.csv  every launch of the eager bench command with its device time (cold caches, serialised:
#                           compare SHARES, not absolutes)
#   <tag>_conv_traffic.csv  DRAM bytes, duration and tensor-pipe activity of EVERY conv / wgrad tcgen05 launch
#   <tag>_prof.ncu-rep      `--set full` capture (source page, stall reasons) of three conv3x3_tc_kernel launches
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained"
export B200_NO_CUDA_GRAPH=1
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
    $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"conv3x3_tc|wgrad3x3_tc|conv_gemm|wgrad_small" -c 420 --csv \
    --log-file gpurun_out/${TAG}_conv_traffic.csv $CMD > gpurun_out/${TAG}_conv_traffic.log 2>&1
echo "traffic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 6 -c 3 -o gpurun_out/${TAG}_prof -f \
    $CMD > gpurun_out/${TAG}_prof.log 2>&1
echo "full rc=$?"
