#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_sr_grad.log
: > $L
timeout 200 python tools/sr_grad_debug.py l1,poisson >> $L 2>&1
timeout 200 python tools/sr_grad_debug.py ms_ssim >> $L 2>&1
XMM_ROW=0 XMM_RDB=0 timeout 200 python tools/sr_grad_debug.py ms_ssim >> $L 2>&1
cat $L
