#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g15.json 2> gpurun_out/r2_g15.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2_g15.json").read().strip().splitlines()[-1])
print("value %.1f e2e 2^20 %.1f 2^22 %.1f 2^24 %.1f"%(l["value"]/1e6,l["e2e"]["value"]/1e6,l["extra"]["e2e_2^22"]["value"]/1e6,l["extra"]["e2e_2^24"]["value"]/1e6))
for k,v in l["extra"].items():
    if "e2e" in v: print(k, round(v["value"]/1e6,2), round(v["e2e"]["value"]/1e6,2), v.get("parity_spot_check"))
    if "e2e_value" in v: print(k, round(v["e2e_value"]/1e6,2), v.get("parity_spot_check"))
PY
tail -2 gpurun_out/r2_g15.err
ECCB200_PIPE_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2> gpurun_out/r2_pipe_trace15.log; grep "eccb200 pipe" gpurun_out/r2_pipe_trace15.log | tail -7
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) 2>&1 | tail -3
