#!/bin/bash
# 8 GPUs, SR training: what does the collective cost?  default (overlapped chunks) vs one serial all-reduce vs none
mkdir -p gpurun_out
L=gpurun_out/r02_scale8_comm_modes.log
: > $L
P=29700
for mode in none overlap serial; do
  XMM_COMM_MODE=$mode XMM_BENCH_WATCHDOG=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --workload train_sr --steps 20 --warmup 3 --no-cpu-baseline --no-train-extra > gpurun_out/r02_s8_$mode.json 2> gpurun_out/r02_s8_$mode.err
  P=$((P+1))
  tail -1 gpurun_out/r02_s8_$mode.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train_sr 8 GPUs XMM_COMM_MODE=$mode', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],3), 'sm_mhz', d['clocks']['sm_mhz'])
except Exception as e: print('$mode FAILED', e)" >> $L
done
cat $L
