for LIB in libupc_b200_b64m9.so libupc_b200_b64m8.so libupc_b200_b32m20.so; do
  for L in 4 8 16; do
    UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/$LIB python bench.py --steps 5 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/tune2_${LIB}_l$L.json 2>> gpurun_out/tune2.err
  done
done
