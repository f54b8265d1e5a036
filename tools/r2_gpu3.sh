#!/bin/bash
# round 2, third GPU call (1 GPU): everything after the safegcd inversion / projective-key verify / drop-in rework
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu3.log 2>&1
tail -8 gpurun_out/r2_pytest_gpu3.log
timeout 900 oracle/_ref/dropin_harness bench libecc_b200/libecc_b200_dropin.so FRP256V1 1048576 > gpurun_out/r2_dropin_bench.log 2>&1; tail -6 gpurun_out/r2_dropin_bench.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_default3.json 2> gpurun_out/r2_bench_default3.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2_bench_default3.json").read().strip().splitlines()[-1])
print("value %.1f M/s ms/step %.3f k1 %.3f k4 %s e2e %.1f (2^22 %.1f, 2^24 %.1f)"%(l["value"]/1e6,l["ms_per_step"],l["roofline"]["kernel_ms"],l["roofline"].get("normalisation_kernel_ms"),l["e2e"]["value"]/1e6,l["extra"]["e2e_2^22"]["value"]/1e6,l["extra"]["e2e_2^24"]["value"]/1e6))
for k,v in l["extra"].items():
    if "value" in v and "kernel_ms" in v: print(k, "%.2f M/s kernel %.3f ms k4 %s e2e %.2f parity %s cpu %s"%(v["value"]/1e6,v["kernel_ms"],v.get("normalisation_ms"),v["e2e"]["value"]/1e6,v["parity_spot_check"],v.get("parity_on_cpu_prefix")))
PY
tail -3 gpurun_out/r2_bench_default3.err
