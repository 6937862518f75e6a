#!/bin/bash
# multi-GPU check (gpurun --gpus N): library-side sharding tests, then the bench at 1, 2, .. N GPUs back to back (strong scaling over C5's edges)
N=${1:-2}
TAG=${2:-r02q}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus_$N.txt 2>&1
[ -n "$SKIP_TESTS" ] || timeout 300 python -m pytest tests/test_cpp_shim.py tests/test_gpu_seeded.py -x -q -m gpu -k "sharded or pipelined" > gpurun_out/${TAG}_tests_$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_tests_$N.log
tail -3 gpurun_out/${TAG}_tests_$N.log
for n in ${NLIST:-1 2 4 8}; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-also --no-plugin --no-cpu-baseline > gpurun_out/${TAG}_bench_${N}box_1.json 2> gpurun_out/${TAG}_bench_${N}box_1.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 \
        > gpurun_out/${TAG}_bench_${N}box_$n.json 2> gpurun_out/${TAG}_bench_${N}box_$n.err
    fi
    echo "bench $n exit $?"
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${N}box_$n.json").read().strip().splitlines()[-1])
    print("n=$n value %.4g ms/step %.2f sim %.2f cluster %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["sim"]["ms_per_step"], d["clustering"]["ms_per_step"], d["e2e"]["value"]))
except Exception as exc:
    print("n=$n no line:", exc)
PY
  fi
done
