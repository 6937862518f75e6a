for LIB in libupc_b200_b64m5.so libupc_b200_b32m9.so; do
  for L in 4; do
    UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/$LIB python bench.py --steps 5 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/tune2_${LIB}_l$L.json 2>> gpurun_out/tune2.err
  done
done
