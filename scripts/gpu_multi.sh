set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/multi_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_1gpu.json 2>> gpurun_out/bench_2gpu.err
