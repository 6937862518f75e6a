#!/bin/bash
# parity tests (bounded), the simulation probe and the default bench line of the in-tree build
# usage: scripts/gpu_r02_check.sh TAG [quick]
TAG=${1:-r02k}
mkdir -p gpurun_out
# the lane sweep first and on its own short leash: a barrier mistake shows up here as a hang, not as a wrong number
timeout 240 python -m pytest tests/test_gpu_parity.py -k "every_lane_count" -x -q -m gpu > gpurun_out/${TAG}_lanes.log 2>&1
rc=$?
echo "lane sweep exit $rc"; tail -2 gpurun_out/${TAG}_lanes.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python scripts/gpu_sim_probe.py seeded 2>&1 | grep -v "^{" > gpurun_out/${TAG}_probe.txt
cat gpurun_out/${TAG}_probe.txt | sed 's/"checksum.*//'
if [ "$2" != "quick" ]; then
  timeout 1000 python -m pytest tests/test_gpu_parity.py tests/test_gpu_seeded.py tests/test_cpp_shim.py -x -q -m gpu > gpurun_out/${TAG}_gputests.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_gputests.log
  tail -3 gpurun_out/${TAG}_gputests.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
echo "bench exit $?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_c5.json'))
print('value %.4g ms/step %.2f sim %.2f cluster %.2f e2e %.4g plugin %s frac %.3f' % (d['value'], d['ms_per_step'], d['sim']['ms_per_step'], d['clustering']['ms_per_step'], d['e2e']['value'], (d.get('e2e_plugin') or {}).get('ms_per_step'), d['roofline']['frac']))
print({k:(v.get('ms_per_step'), v.get('sim_ms'), v.get('clustering_ms')) for k,v in (d.get('also') or {}).items()})"
