#!/bin/bash
# Round profile bundle (run on the GPU box through gpurun; outputs stay small -- gpurun_out/ is capped at 64 MiB):
#   bash tools/profile_round.sh r01_v7 launches   plain bench JSON line, ncu launch list of the same command, conv/HBM tables
#   bash tools/profile_round.sh r01_v7 traffic    DRAM bytes + tensor-pipe % of every conv/wgrad launch of one step (light metric set)
#   bash tools/profile_round.sh r01_v7 full       `ncu --set full` of 24 consecutive conv3x3_tc/wgrad3x3_tc launches of one step
#   bash tools/profile_round.sh r01_v7 widen      the rows next to the hot path: single- vs two-stream step, BASELINE configs 4 / 5,
#                                                 patch-pipeline / eval-metric kernel bandwidth, DP check on one GPU (gloo)
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
elif [ "$what" = "widen" ]; then
  B200_OVERLAP_WGRAD=0 python bench.py --no-cpu-baseline > $out/${tag}_bench_c2_n1_single_stream.json 2> /dev/null
  python bench.py --no-cpu-baseline > $out/${tag}_bench_c2_n1.json 2> /dev/null
  python tools/config_sweep.py c4 > $out/${tag}_c4_base32.json 2> /dev/null
  python tools/config_sweep.py c5 > $out/${tag}_c5.json 2> /dev/null
  python tools/pipeline_bench.py > $out/${tag}_pipeline_bench.jsonl 2> /dev/null
  B200_BUCKET_MB=4 B200_DP_SHARD=0 B200_DP_CHECK_ONE_GPU=1 timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 \
      --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dp_check.py > $out/${tag}_dp_check_one_gpu_gloo.log 2>&1
  python -c "
import json
for f in ('bench_c2_n1_single_stream', 'bench_c2_n1'):
    d = json.load(open('$out/${tag}_%s.json' % f)); print(f, round(d['value']), d['ms_per_step'])
for f in ('c4_base32', 'c5'):
    d = json.load(open('$out/${tag}_%s.json' % f)); print(f, {k: v for k, v in d.items() if k in ('ms_per_step', 'ms_per_batch', 'images_per_s')})
"; tail -1 $out/${tag}_dp_check_one_gpu_gloo.log | cut -c1-200
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
