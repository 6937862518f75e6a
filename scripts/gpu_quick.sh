set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/quick_pytest.log
python bench.py --steps 10 --warmup 3 --lanes 4 > gpurun_out/quick_c2.json 2> gpurun_out/quick.err
python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/quick_c3.json 2>> gpurun_out/quick.err
python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_c1.json 2>> gpurun_out/quick.err
