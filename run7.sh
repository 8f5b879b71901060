mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t_k4.log 2>&1; echo "k4 rc=$?"; tail -n 6 gpurun_out/t_k4.log
timeout -s KILL 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x > gpurun_out/t_m4.log 2>&1; echo "m4 rc=$?"; tail -n 4 gpurun_out/t_m4.log
timeout -s KILL 600 python tools/conv_table.py c2 > gpurun_out/conv_table_c2.log 2>&1; echo "rc=$?"; grep "^{'hw'" gpurun_out/conv_table_c2.log | head -5 | cut -c1-330
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err; cat gpurun_out/bench.log
