#!/bin/bash
# ncu launch list (gpu__time_duration) of the default bench command, short form
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also --no-plugin > gpurun_out/r02e_plain.json 2> gpurun_out/r02e_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches_c5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also --no-plugin > gpurun_out/r02e_ncu_bench.log 2>&1
echo "ncu exit $?"
