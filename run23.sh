mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -n 12 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
