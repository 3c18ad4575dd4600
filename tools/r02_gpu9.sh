#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_train_dbg.log
: > $L
run() { echo "== $*" >> $L; env "$@" XMM_BENCH_WATCHDOG=100 timeout 150 python bench.py --workload train_dn --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -c 700 >> $L; echo >> $L; }
run A=0
run XMM_ROW=0
run XMM_RDB=0
echo "== train_sr default" >> $L
XMM_BENCH_WATCHDOG=100 timeout 150 python bench.py --workload train_sr --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -c 700 >> $L
cat $L
