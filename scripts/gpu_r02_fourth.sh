#!/bin/bash
# round 2, fourth GPU call: A/B of the three sim-kernel changes (variant libraries), parity tests, bench
mkdir -p gpurun_out
P=$PWD/uncertainty_planning_core_b200
python -m pytest tests/test_gpu_parity.py tests/test_gpu_seeded.py -x -q -m gpu -k "not 65536 and not full_size" > gpurun_out/r02d_gputests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02d_gputests.log
tail -3 gpurun_out/r02d_gputests.log
for v in "" _noquad _noresprecull _noplanar; do
  UPC_B200_LIBRARY=$P/libupc_b200$v.so python scripts/gpu_sim_probe.py seeded > gpurun_out/r02d_probe$v.txt 2>&1
  echo "variant '$v':"; tail -1 gpurun_out/r02d_probe$v.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())['result']
print({k:round(v['seeded']['ms'],3) for k,v in d.items()})"
done
python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench_c5.json 2> gpurun_out/r02d_bench_c5.err
echo "bench exit $?"
python -c "
import json
d=json.load(open('gpurun_out/r02d_bench_c5.json'))
print(d['value'], d['ms_per_step'], d['sim']['ms_per_step'], d['clustering']['ms_per_step'], d['e2e']['value'], d['e2e_plugin'].get('ms_per_step'))
print({k:(v.get('ms_per_step'), v.get('sim_ms')) for k,v in d['also'].items()})"
