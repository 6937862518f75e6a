#!/bin/bash
# round 2, first GPU call: bench (C5 default) -> GPU tests -> ncu launch list of the bench command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
echo "bench exit $?" >> gpurun_out/r02_bench_c5.err
tail -c 3000 gpurun_out/r02_bench_c5.json
python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_gputests.log
tail -15 gpurun_out/r02_gputests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c5_reference.json 2>> gpurun_out/r02_bench_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also --no-plugin > gpurun_out/r02_ncu_bench.log 2>&1
echo "ncu exit $?"
