#!/bin/bash
# A/B of an environment switch on the in-tree library: scripts/gpu_ab_env.sh "c1 c2 c3" VAR "valueA valueB"  ("-" = unset)
WORKLOADS=${1:-"c2"}; VAR=$2; VALUES=$3
for rep in 1 2; do
for W in $WORKLOADS; do
  for V in $VALUES; do
    unset $VAR; if [ "$V" != "-" ]; then export $VAR=$V; fi
    S=20; if [ $W != c2 ]; then S=3; fi
    timeout 600 python bench.py --workload $W --steps $S --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$W $VAR=$V', 'ms/step %.3f' % d['ms_per_step'], 'sim kernel ms %.3f' % d['roofline']['kernel_ms'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'big batch %.3g p-steps/s' % d['roofline'].get('big_batch', {}).get('particle_steps_per_s', 0.0))"
  done
done
done
