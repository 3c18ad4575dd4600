#!/bin/bash
# round-2 GPU session 2: fused dense-block kernels -- parity, then timing (every step under its own timeout)
mkdir -p gpurun_out
L=gpurun_out/r02_rdb.log
: > $L
timeout 300 python -m pytest tests/test_gpu_forward.py -x -q -k "fused_dense_block" 2>&1 | tail -25 >> $L
for m in 259 3 2; do
  timeout 90 python tools/rdb_probe.py $m 64 5 >> $L 2>&1 || echo "mode $m batch 64: FAILED/TIMEOUT rc=$?" >> $L
done
XMM_ROW=0 timeout 90 python tools/rdb_probe.py 2 64 5 >> $L 2>&1
for m in 259 2; do
  timeout 90 python tools/rdb_probe.py $m 16 5 >> $L 2>&1 || echo "mode $m batch 16: FAILED/TIMEOUT rc=$?" >> $L
  timeout 90 python tools/rdb_probe.py $m 1 20 >> $L 2>&1 || echo "mode $m batch 1: FAILED/TIMEOUT rc=$?" >> $L
done
cat $L
