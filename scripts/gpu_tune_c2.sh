set -x
python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -5 > gpurun_out/tune_pytest.log
for LIB in libupc_b200.so libupc_b200_b128m3.so libupc_b200_b128m4.so; do
  for L in 2 4 8 16 32; do
    UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/$LIB python bench.py --steps 5 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/tune_${LIB}_l$L.json 2>> gpurun_out/tune.err
  done
done
UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_b128m3.so python scripts/gpu_breakdown.py > gpurun_out/breakdown2.log 2>&1
