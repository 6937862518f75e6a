#!/bin/bash
# old-vs-new library on the secondary workloads (C1, C3) and the C2 breakdown; needs uncertainty_planning_core_b200/libupc_b200_old.so
for W in c1 c3; do
  for L in old new; do
    if [ $L = old ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_old.so; else unset UPC_B200_LIBRARY; fi
    timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$W $L', 'ms/step %.3f' % d['ms_per_step'], 'sim kernel ms %.3f' % d['roofline']['kernel_ms'])"
  done
done
unset UPC_B200_LIBRARY
python scripts/gpu_breakdown.py
