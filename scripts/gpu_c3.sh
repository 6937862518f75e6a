python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c3 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/c3.json 2> gpurun_out/c3.err
for L in 1 2 4 8; do timeout 300 python bench.py --workload c3 --steps 2 --warmup 2 --lanes $L --no-cpu-baseline > gpurun_out/c3_l$L.json 2>> gpurun_out/c3.err; done
timeout 300 python bench.py --workload c3 --impl reference --steps 1 --warmup 0 > gpurun_out/c3_ref.json 2>> gpurun_out/c3.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/c2.json 2>> gpurun_out/c3.err
