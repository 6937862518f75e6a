#!/bin/bash
# round 2, fifth GPU call: clustering rework (fused propagation, merge-proportional exact rounds) - parity tests, bench, launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_seeded.py tests/test_cpp_shim.py -x -q -m gpu > gpurun_out/r02f_gputests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02f_gputests.log
tail -3 gpurun_out/r02f_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench_c5.json 2> gpurun_out/r02f_bench_c5.err
echo "bench exit $?"
python -c "
import json
d=json.load(open('gpurun_out/r02f_bench_c5.json'))
print(d['value'], d['ms_per_step'], d['sim']['ms_per_step'], d['clustering']['ms_per_step'], d['e2e']['value'], d['e2e_plugin'].get('ms_per_step'))
print({k:(v.get('ms_per_step'), v.get('sim_ms'), v.get('clustering_ms')) for k,v in d['also'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_c5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also --no-plugin > gpurun_out/r02f_ncu_bench.log 2>&1
echo "ncu exit $?"
