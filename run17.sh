mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "splitk" > gpurun_out/t_k9.log 2>&1; echo "k9 rc=$?"; tail -n 30 gpurun_out/t_k9.log
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t_k9b.log 2>&1; echo "k9b rc=$?"; tail -n 5 gpurun_out/t_k9b.log
timeout -s KILL 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x > gpurun_out/t_m9.log 2>&1; echo "m9 rc=$?"; tail -n 30 gpurun_out/t_m9.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.log')); print({k:d[k] for k in ('value','ms_per_step','e2e')}); print(d['roofline']); print(d['breakdown_ms'])"
