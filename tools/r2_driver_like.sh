#!/bin/bash
# what the driver does at round end on one GPU: GPU tests, smoke(), reference arm, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_driver_pytest.log 2>&1; tail -4 gpurun_out/r2_driver_pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 ) 2>&1 | cut -c1-300 | tail -5
( time python bench.py --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/r2_driver_bench.json 2> gpurun_out/r2_driver_bench.err; tail -4 gpurun_out/r2_driver_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2_driver_bench.json").read().strip().splitlines()[-1])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in l.items() if k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","gpu_launches","parity_spot_check","parity_on_cpu_prefix")})
print("e2e", l["e2e"]["value"], "roofline", l["roofline"]["frac"], l["roofline"]["frac_executed_imad_wide"], "cpu", l["cpu_baseline"]["value"])
print("extra keys", list(l["extra"].keys()))
PY
