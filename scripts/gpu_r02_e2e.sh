#!/bin/bash
TAG=${1:-r02e2e}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_seeded.py -k "submit_collect or planner_batch" -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/${TAG}_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.4g ms/step %.2f e2e %.4g (%.2f ms; one call %.2f ms) same %s plugin %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_one_call'], d['e2e']['same_outcome_as_device_pass'], (d.get('e2e_plugin') or {}).get('ms_per_step')))"
