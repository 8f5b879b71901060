mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv" > gpurun_out/t_k13.log 2>&1; echo "rc=$?"; tail -n 4 gpurun_out/t_k13.log
for g in 1 2; do B200_CONV_EGROUPS=$g python tools/epi_debug.py 2>&1 | grep "debug=0"; done
timeout -s KILL 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.log')); print({k:d[k] for k in ('value','ms_per_step','e2e')}); print(d['roofline']['frac']); print(d['breakdown_ms'])"
