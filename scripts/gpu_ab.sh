#!/bin/bash
# A/B of two builds of the library on C1 / C2 / C3 (old = uncertainty_planning_core_b200/libupc_b200_old.so, new = the in-tree build)
for W in c1 c3; do
  for L in old new; do
    unset UPC_B200_LIBRARY
    if [ $L = old ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_old.so; fi
    timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$W $L', 'ms/step %.3f' % d['ms_per_step'], 'sim kernel ms %.3f' % d['roofline']['kernel_ms'])"
  done
done
for i in 1 2; do
  for L in old new; do
    unset UPC_B200_LIBRARY
    if [ $L = old ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_old.so; fi
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 $L', 'ms/step %.3f' % d['ms_per_step'], 'sim kernel ms %.3f' % d['roofline']['kernel_ms'], 'e2e ms %.3f' % d['e2e']['ms_per_step'])"
  done
done
