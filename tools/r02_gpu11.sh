#!/bin/bash
# ncu evidence for the CUDA-core / HBM-bound kernels (training step of configs[3] + the inference feed)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_fma.sum,launch__grid_size,launch__block_size
K='regex:normalize|prepare_counts|conv_first|conv_last|loss_reduce|loss_grad|ssim_|msssim|adam|avgpool|colsum|edge_wgrad|restretch|image_upsample|pack_jobs|wgrad_reduce|count_ties|max_kernel|denormalize'
timeout 400 ncu --metrics $M --clock-control none -k "$K" --launch-skip 60 -c 70 --csv --log-file gpurun_out/r02_ncu_hbm_train_sr.csv python bench.py --workload train_sr --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hbm_train.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k "$K" --launch-skip 4 -c 12 --csv --log-file gpurun_out/r02_ncu_hbm_infer.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-extra --no-parity > gpurun_out/ncu_hbm_infer.log 2>&1
wc -l gpurun_out/r02_ncu_hbm_*.csv
