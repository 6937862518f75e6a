#!/bin/bash
# A/B of library builds on the simulation probe (C5 quarter batch, C2, C3): "base" = in-tree libupc_b200.so, V = libupc_b200_V.so
# usage: scripts/gpu_r02_ab.sh TAG "base nogroup ..." [pytest selection]
TAG=${1:-r02h}
VARIANTS=${2:-"base"}
mkdir -p gpurun_out
if [ -n "$3" ]; then
  python -m pytest $3 -x -q -m gpu > gpurun_out/${TAG}_gputests.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_gputests.log
  tail -3 gpurun_out/${TAG}_gputests.log
fi
for rep in ${REPS:-1 2}; do
for L in $VARIANTS; do
  unset UPC_B200_LIBRARY
  if [ $L != base ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_$L.so; fi
  echo "== $L (rep $rep)" >> gpurun_out/${TAG}_ab.txt
  timeout 900 python scripts/gpu_sim_probe.py seeded 2>&1 | grep -v "^{" >> gpurun_out/${TAG}_ab.txt
done
done
cat gpurun_out/${TAG}_ab.txt
