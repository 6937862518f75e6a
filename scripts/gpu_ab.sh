#!/bin/bash
# A/B of builds of the library: "base" = the in-tree libupc_b200.so, any other name V = uncertainty_planning_core_b200/libupc_b200_V.so
# usage: scripts/gpu_ab.sh "c1 c2 c3" "base r4 ..."   (workloads, variants)
WORKLOADS=${1:-"c2"}
VARIANTS=${2:-"base"}
for rep in 1 2; do
for W in $WORKLOADS; do
  for L in $VARIANTS; do
    unset UPC_B200_LIBRARY
    if [ $L != base ]; then export UPC_B200_LIBRARY=$PWD/uncertainty_planning_core_b200/libupc_b200_$L.so; fi
    S=20; if [ $W != c2 ]; then S=3; fi
    timeout 600 python bench.py --workload $W --steps $S --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$W $L', 'ms/step %.3f' % d['ms_per_step'], 'sim kernel ms %.3f' % d['roofline']['kernel_ms'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'big batch %.3g p-steps/s' % d['roofline'].get('big_batch', {}).get('particle_steps_per_s', 0.0))"
  done
done
done
