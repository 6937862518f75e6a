#!/bin/bash
for W in c1 c3; do
  for L in old new newnocull; do
    unset UPC_B200_LIBRARY UPC_B200_NO_FP32_CULL
    if [ $L = old ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_old.so; fi
    if [ $L = newnocull ]; then export UPC_B200_NO_FP32_CULL=1; fi
    timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$W $L', 'ms/step %.3f' % d['ms_per_step'], 'sim kernel ms %.3f' % d['roofline']['kernel_ms'], d['config'].get('lanes'))"
  done
done
