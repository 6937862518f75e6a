#!/bin/bash
# round 2, third GPU call: link-level cull - parity tests, bench, A/B (on / off), phases
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_seeded.py tests/test_cpp_shim.py -x -q -m gpu -k "not 65536" > gpurun_out/r02c_gputests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02c_gputests.log
tail -4 gpurun_out/r02c_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_c5.json 2> gpurun_out/r02c_bench_c5.err
echo "bench exit $?"
python scripts/gpu_sim_probe.py ab > gpurun_out/r02c_probe_ab.txt 2>&1
tail -1 gpurun_out/r02c_probe_ab.txt | cut -c1-1500
UPC_B200_NO_LINK_CULL=1 python scripts/gpu_sim_probe.py seeded > gpurun_out/r02c_probe_nolinkcull.txt 2>&1
tail -1 gpurun_out/r02c_probe_nolinkcull.txt | cut -c1-1000
UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_phases.so python scripts/gpu_sim_probe.py phases > gpurun_out/r02c_probe_phases.txt 2>&1
tail -1 gpurun_out/r02c_probe_phases.txt | cut -c1-3000
