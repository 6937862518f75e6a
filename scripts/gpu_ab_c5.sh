#!/bin/bash
# C5-shaped batch (1,024 edges x 1,000 particles, SE(2), one lane per particle) on library variants
for rep in 1 2; do
for L in $1; do
  unset UPC_B200_LIBRARY
  if [ $L != base ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_$L.so; fi
  timeout 600 python scripts/gpu_c5.py 1024 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5/1024 $L', 'sim ms %.2f' % d['sim_ms'], '%.3g p-steps/s' % d['particle_steps_per_s'])"
done
done
