mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -25 gpurun_out/bench_n2.err | cut -c1-400; head -c 600 gpurun_out/bench_n2.log
