# 8-GPU call: the peer-memory exchange -- correctness, the bench line, the step breakdowns
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
export B200_DP_P2P=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -s KILL 120 $TR --master-port 29621 tools/dp_check.py > gpurun_out/r02_dp_check_n8_p2p.log 2>&1; rc=$?; echo "dp_check rc=$rc"
[ $rc -ne 0 ] && { echo "dp_check failed: not running the rest on possibly wedged GPUs"; exit $rc; }
grep -a "DP check\|Error\|peer exchange" gpurun_out/r02_dp_check_n8_p2p.log | tail -5
timeout -s KILL 300 $TR --master-port 29622 bench.py --gpus 8 > gpurun_out/r02_bench_n8_p2p.json 2> gpurun_out/r02_bench_n8_p2p.err; echo "bench rc=$?"
for c in c2 c3; do
timeout -s KILL 240 $TR --master-port 29623 tools/dp_breakdown.py $c > gpurun_out/r02_dp_breakdown_${c}_n8_p2p.json 2> gpurun_out/r02_dp_breakdown_${c}_n8_p2p.err; echo "bd $c rc=$?"
done
