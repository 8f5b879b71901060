mkdir -p gpurun_out
timeout -s KILL 600 python tools/conv_table.py c2 2>&1 | grep "^hw\|Error\|error" > gpurun_out/conv_table_c2.log; cat gpurun_out/conv_table_c2.log
B200_CONV_GEMM=0 timeout -s KILL 600 python tools/conv_table.py c2 2>&1 | grep "^hw\|Error\|error" | tail -8
