#!/bin/bash
# round 2 evidence: one `ncu --set full` capture per simulation kernel (C5 quarter batch, C2, C3: launches 3 / 7 / 11 of the probe)
# and of the clustering kernels of the 65,536-particle sweep points; raw metric pages and SASS pages are exported on the box.
mkdir -p gpurun_out
TAG=${1:-r02g}
python scripts/gpu_sim_probe.py seeded > gpurun_out/${TAG}_probe.txt 2>&1
echo "probe exit $?"
i=0
for name in c5 c2 c3; do
  skip=$((3 + 4 * i)); i=$((i + 1))
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:forward_simulate --launch-skip $skip --launch-count 1 \
      -o gpurun_out/${TAG}_sim_${name} -f python scripts/gpu_sim_probe.py seeded > gpurun_out/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name exit $?"
  ncu -i gpurun_out/${TAG}_sim_${name}.ncu-rep --page raw --csv > gpurun_out/${TAG}_sim_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_sim_${name}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sim_${name}_sass.csv 2>/dev/null
  ls -la gpurun_out/${TAG}_sim_${name}.ncu-rep
  # the reports carry the whole SASS + source; keep them only when small
  if [ $(stat -c %s gpurun_out/${TAG}_sim_${name}.ncu-rep) -gt 12000000 ]; then rm gpurun_out/${TAG}_sim_${name}.ncu-rep; fi
done
# clustering half: the 65,536-particle points of the sweep, adjacency pass forced and pivot pass
python scripts/gpu_cluster_big.py > gpurun_out/${TAG}_cluster_big.txt 2>&1
echo "cluster big exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"adjacency_kernel|pivot_pass_kernel|root_check_kernel|label_kernel|mirror_lower" --launch-count 40 \
    -o gpurun_out/${TAG}_cluster_big -f python scripts/gpu_cluster_big.py once > gpurun_out/${TAG}_ncu_cluster.log 2>&1
echo "ncu cluster exit $?"
ncu -i gpurun_out/${TAG}_cluster_big.ncu-rep --page raw --csv > gpurun_out/${TAG}_cluster_big_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_cluster_big.ncu-rep
if [ $(stat -c %s gpurun_out/${TAG}_cluster_big.ncu-rep) -gt 12000000 ]; then rm gpurun_out/${TAG}_cluster_big.ncu-rep; fi
