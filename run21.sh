mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t_k11.log 2>&1; echo "k11 rc=$?"; tail -n 8 gpurun_out/t_k11.log
timeout -s KILL 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x > gpurun_out/t_m11.log 2>&1; echo "m11 rc=$?"; tail -n 8 gpurun_out/t_m11.log
timeout -s KILL 300 python tools/mem_table.py > gpurun_out/mem_table.log 2>&1; echo "mem rc=$?"; tail -8 gpurun_out/mem_table.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.log')); print({k:d[k] for k in ('value','ms_per_step','e2e')}); print(d['roofline']); print(d['breakdown_ms'])"
