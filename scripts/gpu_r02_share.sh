#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -k "every_lane_count or cull_changes_nothing" -x -q -m gpu > gpurun_out/${TAG}_lanes.log 2>&1
echo "lane sweep exit $?"; tail -2 gpurun_out/${TAG}_lanes.log
for L in base noshare base noshare; do
  unset UPC_B200_LIBRARY
  if [ $L != base ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_$L.so; fi
  echo "== $L" >> gpurun_out/${TAG}_ab.txt
  timeout 300 python scripts/gpu_sim_probe.py seeded 2>&1 | grep -v "^{" >> gpurun_out/${TAG}_ab.txt
done
cat gpurun_out/${TAG}_ab.txt | sed 's/"checksum.*//'
