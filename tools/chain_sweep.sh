#!/bin/bash
# Developer tool: sweep the SM split / segment count of the pipelined dense-block kernel (run under gpurun).
for split in 24,25,27,31,41 25,26,28,30,39 23,24,27,32,42 26,26,27,29,40 24,24,26,30,44; do
  for segs in 2; do
  echo "== segs=$segs split=$split"; XMM_CHAIN_SEGS=$segs XMM_CHAIN_SPLIT=$split XMM_CHAIN_PROF=1 timeout 60 build/probe sweep 2>&1 | tail -7 | grep "chain layer\|mode=1"
  done
done
echo "== segs=1 split=24,25,27,31,41"; XMM_CHAIN_SEGS=1 XMM_CHAIN_SPLIT=24,25,27,31,41 timeout 60 build/probe sweep 2>&1 | grep "mode=1"
echo "== segs=3 split=24,25,27,31,41"; XMM_CHAIN_SEGS=3 XMM_CHAIN_SPLIT=24,25,27,31,41 timeout 60 build/probe sweep 2>&1 | grep "mode=1"
echo "== B=64 split=24,25,27,31,41"; XMM_CHAIN_SPLIT=24,25,27,31,41 timeout 60 build/probe sweep 64 2>&1 | grep "time chain"
