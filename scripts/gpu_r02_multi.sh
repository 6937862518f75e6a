#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): library-side sharding tests, then the bench at 1, 2, 4, ... N GPUs back to back
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02m_gpus_$N.txt 2>&1
python -m pytest tests/test_cpp_shim.py tests/test_gpu_seeded.py -x -q -m gpu -k "sharded" > gpurun_out/r02m_tests_$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02m_tests_$N.log
tail -3 gpurun_out/r02m_tests_$N.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02m_reference_$N.json 2> gpurun_out/r02m_reference_$N.err
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps 10 --warmup 3 --no-also > gpurun_out/r02m_bench_${N}box_1.json 2> gpurun_out/r02m_bench_${N}box_1.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 \
        > gpurun_out/r02m_bench_${N}box_$n.json 2> gpurun_out/r02m_bench_${N}box_$n.err
    fi
    echo "bench $n exit $?"
    tail -c 400 gpurun_out/r02m_bench_${N}box_$n.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02m_bench_${N}box_$n.json").read().strip().splitlines()[-1])
    print("n=$n value %.4g ms/step %.2f sim %.2f cluster %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["sim"]["ms_per_step"], d["clustering"]["ms_per_step"], d["e2e"]["value"]))
except Exception as exc:
    print("n=$n no line:", exc)
PY
  fi
done
if [ $N -gt 1 ]; then
  python - <<PY > gpurun_out/r02m_plugin_sharded_$N.json 2> gpurun_out/r02m_plugin_sharded_$N.err
import json, sys
sys.path.insert(0, ".")
import bench
wl = bench.make_workload("c5")
cp = bench.cluster_params()
out = {}
for g in (1, $N):
    out[str(g)] = bench.run_plugin_driver(wl, cp, "edges", 4, 1, gpus=g)
print(json.dumps(out))
PY
  cat gpurun_out/r02m_plugin_sharded_$N.json | cut -c1-1200
fi
