# one 8-GPU call: DP correctness, the bench line, the step breakdowns (default and variants)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -s KILL 240 $TR --master-port 29601 tools/dp_check.py > gpurun_out/r02_dp_check_n8_head.log 2>&1; echo "dp_check rc=$?"
timeout -s KILL 300 $TR --master-port 29602 bench.py --gpus 8 > gpurun_out/r02_bench_n8_head.json 2> gpurun_out/r02_bench_n8_head.err; echo "bench rc=$?"
for c in c2 c3; do
timeout -s KILL 240 $TR --master-port 29603 tools/dp_breakdown.py $c > gpurun_out/r02_dp_breakdown_${c}_n8_head.json 2> gpurun_out/r02_dp_breakdown_${c}_n8_head.err; echo "bd $c rc=$?"
done
B200_HEAD_BUCKET_MB=0 timeout -s KILL 240 $TR --master-port 29604 tools/dp_breakdown.py c2 > gpurun_out/r02_dp_breakdown_c2_n8_nohead.json 2> gpurun_out/r02_dp_breakdown_c2_n8_nohead.err; echo "bd c2 nohead rc=$?"
NCCL_MAX_CTAS=8 timeout -s KILL 240 $TR --master-port 29605 tools/dp_breakdown.py c3 > gpurun_out/r02_dp_breakdown_c3_n8_ctas8.json 2> gpurun_out/r02_dp_breakdown_c3_n8_ctas8.err; echo "bd c3 ctas8 rc=$?"
B200_BUCKET_MB=128 timeout -s KILL 240 $TR --master-port 29606 tools/dp_breakdown.py c3 > gpurun_out/r02_dp_breakdown_c3_n8_b128.json 2> gpurun_out/r02_dp_breakdown_c3_n8_b128.err; echo "bd c3 b128 rc=$?"
tail -3 gpurun_out/r02_dp_check_n8_head.log
