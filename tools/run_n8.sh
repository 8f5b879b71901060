mkdir -p gpurun_out
for c in c3 c2; do
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$( [ $c = c3 ] && echo 3 || echo 4 ) bench.py --gpus 8 --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${c}_n8.log 2> gpurun_out/bench_${c}_n8.err; echo "$c n8 rc=$?"
done
sleep 1
for c in c3 c2; do python -c "
import json; d=json.load(open('gpurun_out/bench_${c}_n8.log')); print('$c', {k:d[k] for k in ('value','ms_per_step','e2e','n_gpus')}, d['clocks'])"; done
