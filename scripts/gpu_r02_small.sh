#!/bin/bash
# the per-rank share of C5 at 8 GPUs (512 edges) on one GPU: where do the fixed costs of a step go?  + DRAM traffic of one full C5 launch
TAG=${1:-r02r}
mkdir -p gpurun_out
for e in 512 1024 4096; do
  timeout 300 python bench.py --edges $e --steps 10 --warmup 3 --no-also --no-plugin --no-cpu-baseline > gpurun_out/${TAG}_bench_e$e.json 2> gpurun_out/${TAG}_bench_e$e.err
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_e$e.json'))
print('edges $e: step %.3f ms sim %.3f cluster %.3f e2e %.3f ms launches %s' % (d['ms_per_step'], d['sim']['ms_per_step'], d['clustering']['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches']))"
done
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:forward_simulate --launch-skip 2 --launch-count 1 --csv \
   --log-file gpurun_out/${TAG}_traffic_c5.csv python bench.py --steps 2 --warmup 1 --no-also --no-plugin --no-cpu-baseline > gpurun_out/${TAG}_traffic_c5.log 2>&1
tail -3 gpurun_out/${TAG}_traffic_c5.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_e512.csv \
   python bench.py --edges 512 --steps 2 --warmup 1 --no-also --no-plugin --no-cpu-baseline > gpurun_out/${TAG}_launches_e512.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_e512.csv | head -30
