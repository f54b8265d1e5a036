#!/bin/bash
# K6 (multi-scalar-multiplication batch verification): GPU tests, the bench legs, per-kernel times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_msm.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_msm_tests.log
( timeout 900 python -c "
import json, bench
print(json.dumps(bench.run_schnorr_msm(0, 'ecfsdsa')))
print(json.dumps(bench.run_schnorr_msm(0, 'bip0340')))
" 2>&1 | tail -6 ) > gpurun_out/r2_msm_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"k_msm" --csv \
  --log-file gpurun_out/r2_msm_launches.csv python tools/msm_once.py > gpurun_out/r2_msm_ncu.log 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_msm_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:][-14:]:
    print(r[ki][:60], r[vi])
P
cat gpurun_out/r2_msm_tests.log gpurun_out/r2_msm_bench.log; tail -3 gpurun_out/r2_msm_ncu.log
