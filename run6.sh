mkdir -p gpurun_out
export B200_NO_CUDA_GRAPH=1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launchlist rc=$?"
python tools/profile_kernels.py > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 21 -c 21 -o gpurun_out/prof_r01 python tools/profile_kernels.py > gpurun_out/ncu_prof.log 2>&1
echo "full rc=$?"; tail -n 3 gpurun_out/ncu_prof.log; ls -la gpurun_out | tail -n 8
