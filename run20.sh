mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t_k10.log 2>&1; echo "k10 rc=$?"; tail -n 5 gpurun_out/t_k10.log
B200_NO_CUDA_GRAPH=1 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_bench.log | cut -c1-300
