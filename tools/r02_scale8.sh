#!/bin/bash
# 8-GPU run: training A/B of the comm knobs (NCCL CTA cap, SM reserve), then inference; 1-GPU references on the same box
mkdir -p gpurun_out
L=gpurun_out/r02_scale8.log
: > $L
P=29600
one() {  # name, nproc, workload, env...
  local name=$1 n=$2 w=$3; shift 3
  if [ "$n" = "1" ]; then
    env "$@" XMM_BENCH_WATCHDOG=200 timeout 300 python bench.py --gpus 1 --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-train-extra > gpurun_out/r02_s_${name}.json 2> gpurun_out/r02_s_${name}.err
  else
    env "$@" XMM_BENCH_WATCHDOG=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-train-extra > gpurun_out/r02_s_${name}.json 2> gpurun_out/r02_s_${name}.err
    P=$((P+1))
  fi
  tail -1 gpurun_out/r02_s_${name}.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print('$name FAILED', e)" >> $L
}
one train_sr_1 1 train_sr A=0
one train_dn_1 1 train_dn A=0
one infer_1 1 infer_sr A=0
one train_sr_8_default 8 train_sr A=0
one train_sr_8_nocap 8 train_sr NCCL_MAX_CTAS=32 XMM_COMM_SM_RESERVE=0
one train_sr_8_cap_only 8 train_sr NCCL_MAX_CTAS=4 XMM_COMM_SM_RESERVE=0
one train_sr_8_res8 8 train_sr NCCL_MAX_CTAS=8 XMM_COMM_SM_RESERVE=8
one train_dn_8_default 8 train_dn A=0
one infer_8 8 infer_sr A=0
cat $L
