#!/bin/bash
TAG=${1:-r02p3}
mkdir -p gpurun_out
for e in 512 1024 2048; do
  timeout 300 python bench.py --edges $e --steps 10 --warmup 3 --no-also --no-plugin --no-cpu-baseline > gpurun_out/${TAG}_bench_e$e.json 2> gpurun_out/${TAG}_bench_e$e.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_e$e.json'))
print('edges $e: step %.3f ms (back to back %.3f) sim %.3f cluster %.3f e2e %.3f ms value %.4g' % (d['ms_per_step'], d['pipeline']['ms_per_step_back_to_back'], d['sim']['ms_per_step'], d['clustering']['ms_per_step'], d['e2e']['ms_per_step'], d['value']))"
done
