#!/bin/bash
TAG=${1:-r02pl}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_cpp_shim.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.4g e2e %.4g (%.2f ms) plugin %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], json.dumps(d.get('e2e_plugin'))[:400]))"
