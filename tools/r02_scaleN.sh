#!/bin/bash
# N-GPU weak-scaling point (inference, DN training, SR training; 20 steps each)
N=$1
mkdir -p gpurun_out
L=gpurun_out/r02_scale_${N}gpu.log
: > $L
P=29800
for w in infer_sr train_dn train_sr; do
  XMM_BENCH_WATCHDOG=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-train-extra --no-parity > gpurun_out/r02_bench_scale_${w}_${N}.json 2> gpurun_out/r02_scale_${w}_${N}.err
  P=$((P+1))
  tail -1 gpurun_out/r02_bench_scale_${w}_${N}.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$N GPUs $w', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'sm_mhz', d['clocks']['sm_mhz'])
except Exception as e: print('$w FAILED', e)" >> $L
done
cat $L
