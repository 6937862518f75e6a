# 1 / 2 / 4 / 8 GPUs of one box, back to back (the driver's SCALE run does the same at round end)
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale8_1.json 2> gpurun_out/scale8.err
P=29520
for N in 2 4 8; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale8_$N.json 2>> gpurun_out/scale8.err
done
for N in 1 2 4 8; do tail -1 gpurun_out/scale8_$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($N, 'GPUs: value %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'])"; done
