# Round-1 profile of the default bench command (B200_PROFILING.md recipe): plain run first (must exit 0),
# then the ncu launch list of the same command, then one full capture of the dominant kernel.
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/prof_plain2.json 2>> gpurun_out/prof_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:forward_simulate -s 3 -c 1 -o gpurun_out/prof_sim_c2 -f $CMD > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref_final.json 2>> gpurun_out/bench_c2_final.err
UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_phase.so python scripts/gpu_phases.py > gpurun_out/phases.log 2>&1
python scripts/gpu_breakdown.py > gpurun_out/breakdown3.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
