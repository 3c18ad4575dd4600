python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_training.py -x -q 2>&1 | tail -3
for r in 1 0 1 0; do XMM_WG_ROWS8=$r python tools/wgrad_time.py 2>&1 | grep "stacked=1" | sed "s/^/rows8=$r /"; done | tee gpurun_out/wgrad_rows8.log
for g in 1 0 1 0; do XMM_TRAIN_GRAPH=$g python bench.py --workload train_dn --steps 8 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train_dn graph=$g', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['config'].get('final_loss'))"; done 2>&1 | tee gpurun_out/bench_train_graph_ab.log
XMM_TRAIN_GRAPH=1 python bench.py --workload train_sr --steps 8 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train_sr graph=1', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['config'].get('final_loss'))" | tee -a gpurun_out/bench_train_graph_ab.log
