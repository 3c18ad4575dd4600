#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_sr_grad2.log
: > $L
timeout 200 python tools/sr_grad_debug2.py sr >> $L 2>&1
XMM_ROW=0 XMM_RDB=0 timeout 200 python tools/sr_grad_debug2.py sr >> $L 2>&1
cat $L
