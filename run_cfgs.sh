mkdir -p gpurun_out
for c in c3; do
timeout -s KILL 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; echo "$c rc=$?"; tail -3 gpurun_out/bench_$c.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_$c.log')); print({k:d[k] for k in ('value','ms_per_step','e2e')}); print(d['config']['step_frac_of_conv_roofline'], d['roofline']['frac']); print(d['breakdown_ms'])"
done
