python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; cat gpurun_out/bench_pair.json | head -c 2500; echo
XMM_DX_PAIR=0 python bench.py --steps 10 --warmup 3 --no-train-extra --no-cpu-baseline > gpurun_out/bench_nopair.json 2> gpurun_out/bench_nopair.err; cat gpurun_out/bench_nopair.json | head -c 600; echo
