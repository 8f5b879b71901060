#!/bin/bash
# Round profile bundle (run on the GPU box through gpurun; outputs stay small -- gpurun_out/ is capped at 64 MiB):
#   bash tools/profile_round.sh r01_v7 launches   plain bench JSON line, ncu launch list of the same command, conv/HBM tables
#   bash tools/profile_round.sh r01_v7 traffic    DRAM bytes + tensor-pipe % of every conv/wgrad launch of one step (light metric set)
#   bash tools/profile_round.sh r01_v7 full       `ncu --set full` of 24 consecutive conv3x3_tc/wgrad3x3_tc launches of one step
tag=${1:-r01}
what=${2:-launches}
out=gpurun_out
mkdir -p $out
if [ "$what" = "launches" ]; then
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { echo "bench failed"; tail -5 $out/${tag}_bench.err; exit 1; }
  B200_NO_CUDA_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $out/${tag}_ncu_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1 || echo "ncu launch list failed"
  python tools/conv_table.py c2 2>&1 | grep "^hw" > $out/${tag}_conv_table_c2.txt
  python tools/mem_table.py > $out/${tag}_mem_table.txt 2>&1
  head -c 300 $out/${tag}_bench.json; echo
elif [ "$what" = "traffic" ]; then
  # DRAM bytes / tensor-pipe cycles of EVERY conv3x3_tc / wgrad launch of one step (same population as bench's `achieved`)
  B200_NO_CUDA_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:'conv3x3_tc_kernel|wgrad3x3_tc_kernel|conv_gemm_kernel|wgrad_small_kernel' --launch-skip 213 --launch-count 71 --csv \
      --log-file $out/${tag}_conv_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_conv_traffic.log 2>&1 || echo "ncu traffic failed"
  tail -2 $out/${tag}_conv_traffic.csv | cut -c1-300
else
  B200_NO_CUDA_GRAPH=1 ncu --set full --clock-control none -k regex:'conv3x3_tc_kernel|wgrad3x3_tc_kernel' \
      --launch-skip 118 --launch-count 24 -o /tmp/${tag}_conv_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1 || echo "ncu full failed"
  ncu -i /tmp/${tag}_conv_full.ncu-rep --page raw --csv > $out/${tag}_conv_full_raw.csv 2>/dev/null
  ls -la /tmp/${tag}_conv_full.ncu-rep $out/${tag}_conv_full_raw.csv
  tail -2 $out/${tag}_ncu_full.log | cut -c1-200
fi
