set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/quick_pytest.log
UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_phase.so python scripts/gpu_phases.py > gpurun_out/phases.log 2>&1
for L in 0 2 8; do python bench.py --steps 10 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/quick_c2_l$L.json 2>> gpurun_out/quick.err; done
python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_c1.json 2>> gpurun_out/quick.err
python scripts/gpu_breakdown.py > gpurun_out/breakdown3.log 2>&1
