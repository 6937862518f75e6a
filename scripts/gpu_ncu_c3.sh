CMD="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_c3_plain.json 2> gpurun_out/prof_c3_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:forward_simulate -s 3 -c 1 -o gpurun_out/prof_sim_c3 -f $CMD > gpurun_out/ncu_c3.log 2>&1
