#!/bin/bash
# round 2, second GPU call: cluster/seeded tests after the clustering rework, bench, tape-vs-seeded A/B, phase cycles, ncu --set full
mkdir -p gpurun_out
python -m pytest tests/test_gpu_seeded.py tests/test_gpu_parity.py tests/test_cpp_shim.py -x -q -m gpu > gpurun_out/r02b_gputests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02b_gputests.log
tail -5 gpurun_out/r02b_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err
echo "bench exit $?"
python scripts/gpu_sim_probe.py ab > gpurun_out/r02b_probe_ab.txt 2>&1
tail -4 gpurun_out/r02b_probe_ab.txt | cut -c1-600
UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_phases.so python scripts/gpu_sim_probe.py phases > gpurun_out/r02b_probe_phases.txt 2>&1
tail -4 gpurun_out/r02b_probe_phases.txt | cut -c1-1500
# ncu --set full: one launch of the simulation kernel on a quarter of C5, on C3, and the clustering kernels of C5
ncu --set full --clock-control none --import-source on -k regex:forward_simulate -s 2 -c 1 -o gpurun_out/r02b_ncu_sim_c5 -f \
    python bench.py --edges 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-also --no-plugin > gpurun_out/r02b_ncu_c5.log 2>&1
echo "ncu c5 exit $?"
ncu --set full --clock-control none --import-source on -k regex:forward_simulate -s 1 -c 1 -o gpurun_out/r02b_ncu_sim_c3 -f \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-also --no-plugin > gpurun_out/r02b_ncu_c3.log 2>&1
echo "ncu c3 exit $?"
