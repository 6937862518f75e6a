#!/bin/bash
# simulation-kernel CTA size experiment: same code, 32 / 64 / 128 threads per CTA
for B in 128 32 64 128 32 64; do
  unset UPC_B200_LIBRARY
  if [ $B != 128 ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_b$B.so; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('block $B', 'ms/step %.3f' % d['ms_per_step'], 'sim kernel ms %.3f' % d['roofline']['kernel_ms'])"
done
