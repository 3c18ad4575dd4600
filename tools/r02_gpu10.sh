#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_sr_grad3.log
: > $L
timeout 200 python tools/sr_grad_debug2.py sr >> $L 2>&1
cat $L
