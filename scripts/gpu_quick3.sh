python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/quick_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/q3_c2.json 2> gpurun_out/q3.err
python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q3_c1.json 2>> gpurun_out/q3.err
