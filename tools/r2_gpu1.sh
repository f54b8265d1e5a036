#!/bin/bash
# round 2, first GPU call: parity suite, comb-window sweep, full default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log
for w in 22 24 26; do
  timeout 600 python bench.py --steps 10 --warmup 3 --comb-window $w --no-extra --no-cpu-baseline > gpurun_out/r2_bench_w$w.json 2> gpurun_out/r2_bench_w$w.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r2_bench_w$w.json").read().strip().splitlines()[-1])
    print("w=$w value %.1f M/s e2e %.1f M/s k1 %.3f ms frac_exec %.3f parity %s"%(l["value"]/1e6,l["e2e"]["value"]/1e6,l["roofline"]["kernel_ms"],l["roofline"]["frac_executed_imad_wide"],l["parity_spot_check"]))
except Exception as e:
    print("w=$w failed", e); print(open("gpurun_out/r2_bench_w$w.err").read()[-1500:])
PY
done
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 3000 gpurun_out/r2_bench_default.json; tail -5 gpurun_out/r2_bench_default.err
