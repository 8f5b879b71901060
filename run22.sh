mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_entry_points_gpu.py -q -m gpu -x > gpurun_out/t_e1.log 2>&1; echo "e1 rc=$?"; tail -n 40 gpurun_out/t_e1.log
