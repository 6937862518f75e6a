python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/quick_pytest.log
python scripts/gpu_cluster_sweep.py > gpurun_out/cluster_sweep.log 2>&1
python bench.py --workload c3 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/c3.json 2> gpurun_out/c3.err
