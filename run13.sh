mkdir -p gpurun_out
for v in 0 1; do
B200_LN_BWD_MP=$v B200_RESAMPLE_V=$v timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'ln_bwd|ln_fwd|resample' -o gpurun_out/mem_v$v -f python tools/mem_probe.py > gpurun_out/ncu_mem_v$v.log 2>&1; echo "ncu v$v rc=$?"; tail -n 3 gpurun_out/ncu_mem_v$v.log
ncu -i gpurun_out/mem_v$v.ncu-rep --page raw --csv > gpurun_out/mem_v${v}_raw.csv 2>/dev/null
done
B200_LN_BWD_MP=0 B200_RESAMPLE_V=0 python tools/mem_table.py 2>&1 | head -17
