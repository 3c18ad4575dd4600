#!/bin/bash
# Convenience wrapper used with gpurun: GPU tests + smoke, logs into gpurun_out/.
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -s 2>&1 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
