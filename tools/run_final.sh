set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --workload train_dn --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_train_dn.json 2>> gpurun_out/bench_final.err
python bench.py --workload train_sr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_train_sr.json 2>> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 70 --csv --log-file gpurun_out/launches_infer_b64_v2.csv python bench.py --steps 1 --warmup 3 --no-train-extra --no-cpu-baseline > gpurun_out/ncu_infer.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2450 -c 820 --csv --log-file gpurun_out/launches_train_dn_v2.csv python bench.py --workload train_dn --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 30 -c 3 -f -o gpurun_out/prof_wgrad_v2 python bench.py --workload train_dn --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
