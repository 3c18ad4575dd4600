#!/bin/bash
# round-2 GPU session 1: locate the row-hop hang (every step under its own timeout)
mkdir -p gpurun_out
L=gpurun_out/r02_row_probe.log
: > $L
for c in c32 c64 c96 c128 c160 c5 c5e trunk dg128 hr; do
  for m in 9 0; do
    timeout 60 python tools/row_probe.py $c $m 64 >> $L 2>&1 || echo "$c tap_mode $m batch 64: FAILED/TIMEOUT rc=$?" >> $L
  done
done
XMM_ROW=0 timeout 60 python tools/row_probe.py c160 0 64 >> $L 2>&1
XMM_ROW=0 timeout 60 python tools/row_probe.py c64 0 64 >> $L 2>&1
XMM_ROW=0 timeout 60 python tools/row_probe.py c32 0 64 >> $L 2>&1
for c in c32 c96 c160 c5e dg128; do
  timeout 60 python tools/row_probe.py $c 9 16 >> $L 2>&1 || echo "$c tap_mode 9 batch 16: FAILED/TIMEOUT rc=$?" >> $L
  XMM_ROW=0 timeout 60 python tools/row_probe.py $c 0 16 >> $L 2>&1
done
cat $L
timeout 300 python -m pytest tests/test_gpu_forward.py -x -q -k "row_hop" 2>&1 | tail -5
timeout 400 python -m pytest tests/test_gpu_forward.py -x -q -k "not row_hop" 2>&1 | tail -15
