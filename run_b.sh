mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_golden.py -q -m gpu -x > gpurun_out/t_k12.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/t_k12.log
timeout -s KILL 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.log')); print({k:d[k] for k in ('value','ms_per_step','e2e')}); print(d['roofline']['frac'], d['roofline']['wgrad_kernel']); print(d['breakdown_ms'])"
