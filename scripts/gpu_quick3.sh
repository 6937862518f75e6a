python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/quick_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/q3_c2.json 2> gpurun_out/q3.err
