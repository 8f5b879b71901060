mkdir -p gpurun_out
timeout -s KILL 300 python tools/conv_debug.py > gpurun_out/conv_debug.log 2>&1; echo "rc=$?"; cat gpurun_out/conv_debug.log
