mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp rc=$?"; tail -4 gpurun_out/dp_check.log | cut -c1-300
B200_DP_SHARD=0 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_check.py > gpurun_out/dp_check0.log 2>&1; echo "dp0 rc=$?"; tail -2 gpurun_out/dp_check0.log | cut -c1-300
for c in c2 c3; do
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${c}_n2.log 2> gpurun_out/bench_${c}_n2.err; echo "$c n2 rc=$?"; tail -3 gpurun_out/bench_${c}_n2.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_${c}_n2.log')); print({k:d[k] for k in ('value','ms_per_step','e2e','n_gpus')})"
done
