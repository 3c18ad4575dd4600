#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_rdb_dbg.log
: > $L
timeout 200 python -m pytest tests/test_gpu_forward.py -x -q -k "fused_dense_block" 2>&1 | tail -40 >> $L
for m in 259 3 2; do
  timeout 60 python tools/rdb_probe.py $m 64 5 >> $L 2>&1 || echo "mode $m batch 64: FAILED/TIMEOUT rc=$?" >> $L
done
for m in 259 2; do
  timeout 60 python tools/rdb_probe.py $m 16 5 >> $L 2>&1 || echo "mode $m batch 16: FAILED/TIMEOUT rc=$?" >> $L
  timeout 60 python tools/rdb_probe.py $m 1 20 >> $L 2>&1 || echo "mode $m batch 1: FAILED/TIMEOUT rc=$?" >> $L
done
cat $L
