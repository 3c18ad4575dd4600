N=$1
P=29500
for w in infer_sr train_dn train_sr; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-train-extra > gpurun_out/bench_${N}gpu_$w.json 2> gpurun_out/bench_${N}gpu_$w.err
  tail -1 gpurun_out/bench_${N}gpu_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$N GPUs $w', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
  P=$((P+1))
done
