set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for L in 1 2 4 8 16 32; do python bench.py --steps 5 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_c2_l$L.json 2>> gpurun_out/bench_c2.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_ref.json 2>> gpurun_out/bench_c2.err
python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1.json 2>> gpurun_out/bench_c2.err
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err
