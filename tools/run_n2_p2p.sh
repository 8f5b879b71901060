# 2-GPU call: the peer-memory exchange against NCCL -- correctness (dp_check), then the step breakdown
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B200_DP_P2P=1 timeout -s KILL 200 $TR --master-port 29611 tools/dp_check.py > gpurun_out/r02_dp_check_n2_p2p.log 2>&1; rc=$?; echo "dp_check p2p rc=$rc"
[ $rc -ne 0 ] && { echo "dp_check failed: not running the rest on possibly wedged GPUs"; exit $rc; }
grep -a "DP check\|Error\|error\|peer" gpurun_out/r02_dp_check_n2_p2p.log | tail -8
for c in c2 c3; do
B200_DP_P2P=1 timeout -s KILL 200 $TR --master-port 29612 tools/dp_breakdown.py $c > gpurun_out/r02_dp_breakdown_${c}_n2_p2p.json 2> gpurun_out/r02_dp_breakdown_${c}_n2_p2p.err; echo "bd $c p2p rc=$?"
timeout -s KILL 200 $TR --master-port 29613 tools/dp_breakdown.py $c > gpurun_out/r02_dp_breakdown_${c}_n2_nccl.json 2> gpurun_out/r02_dp_breakdown_${c}_n2_nccl.err; echo "bd $c nccl rc=$?"
done
tail -3 gpurun_out/r02_dp_breakdown_c2_n2_p2p.err
