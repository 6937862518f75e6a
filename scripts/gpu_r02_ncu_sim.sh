#!/bin/bash
# one `ncu --set full` capture per simulation kernel of the probe (C5 quarter batch, C2, C3 = launches 3 / 7 / 11), pages exported on the box
TAG=${1:-r02l}
WHICH=${2:-"c5 c2 c3"}
mkdir -p gpurun_out
for name in $WHICH; do
  case $name in c5) skip=3;; c2) skip=7;; c3) skip=11;; esac
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:forward_simulate --launch-skip $skip --launch-count 1 \
      -o gpurun_out/${TAG}_sim_${name} -f python scripts/gpu_sim_probe.py seeded > gpurun_out/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name exit $?"
  ncu -i gpurun_out/${TAG}_sim_${name}.ncu-rep --page raw --csv > gpurun_out/${TAG}_sim_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_sim_${name}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sim_${name}_sass.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_sim_${name}.ncu-rep
done
