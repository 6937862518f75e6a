#!/bin/bash
# final records of the round: the default bench line, the ncu launch list of the same command (short form), one full capture of the
# dominant kernel (C5 quarter batch of the probe) with its source page
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
echo "bench exit $?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_c5.json'))
print('value %.4g ms/step %.2f (b2b %.2f) sim %.2f cluster %.2f e2e %.4g same %s plugin %s frac %.3f' % (d['value'], d['ms_per_step'], d['pipeline']['ms_per_step_back_to_back'], d['sim']['ms_per_step'], d['clustering']['ms_per_step'], d['e2e']['value'], d['e2e']['same_outcome_as_device_pass'], (d.get('e2e_plugin') or {}).get('ms_per_step'), d['roofline']['frac']))
print({k:(v.get('ms_per_step'), v.get('sim_ms'), v.get('clustering_ms')) for k,v in (d.get('also') or {}).items()})"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also --no-plugin > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
bash scripts/gpu_r02_ncu_sim.sh ${TAG} "c5 c3"
