#!/bin/bash
# pipelined step: parity test of the split entry points, bench at the full batch and at the per-rank share of 8 GPUs (512 edges)
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_seeded.py tests/test_gpu_parity.py -k "pipelined or sharded or planner_batch or clustering_matches_oracle or c5_properties" -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_tests.log
for e in 512 4096; do
  timeout 300 python bench.py --edges $e --steps 10 --warmup 3 --no-also --no-plugin --no-cpu-baseline > gpurun_out/${TAG}_bench_e$e.json 2> gpurun_out/${TAG}_bench_e$e.err
  tail -2 gpurun_out/${TAG}_bench_e$e.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_e$e.json'))
print('edges $e: step %.3f ms (back to back %.3f) sim %.3f cluster %.3f e2e %.3f ms value %.4g' % (d['ms_per_step'], d['pipeline']['ms_per_step_back_to_back'], d['sim']['ms_per_step'], d['clustering']['ms_per_step'], d['e2e']['ms_per_step'], d['value']))"
done
