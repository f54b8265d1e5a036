#!/bin/bash
# K6 (multi-scalar-multiplication batch verification): GPU tests, the drop-in harness (new generic entries), the bench leg
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_msm.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_msm_tests.log
( timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -15 ) >> gpurun_out/r2_msm_tests.log
( timeout 600 python -c "
import json, bench
print(json.dumps(bench.run_ecfsdsa_msm(0)))
" 2>&1 | tail -5 ) > gpurun_out/r2_msm_bench.log
cat gpurun_out/r2_msm_tests.log gpurun_out/r2_msm_bench.log
