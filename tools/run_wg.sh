set -x
python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_training.py -x -q > gpurun_out/pytest_wg.log 2>&1; echo rc=$? >> gpurun_out/pytest_wg.log
tail -3 gpurun_out/pytest_wg.log
python bench.py --workload train_dn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tdn.json 2> gpurun_out/bench_tdn.err; tail -c 1500 gpurun_out/bench_tdn.json
python bench.py --workload train_sr --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tsr.json 2> gpurun_out/bench_tsr.err; tail -c 1500 gpurun_out/bench_tsr.json
