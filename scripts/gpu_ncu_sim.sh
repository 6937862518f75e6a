CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain2.json 2> gpurun_out/prof_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:forward_simulate -s 3 -c 1 -o gpurun_out/prof_sim_c2 -f $CMD > gpurun_out/ncu_full.log 2>&1
