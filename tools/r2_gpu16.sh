#!/bin/bash
# session 3 of round 2: GPU tests (incl. the C harness with the five new schemes), the new verify_batch adapters at
# 2^20 real ec_pub_key structs, and the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2s3_pytest.log 2>&1; tail -6 gpurun_out/r2s3_pytest.log
: > gpurun_out/r2s3_dropin_ds_bench.log
for sc in ECGDSA ECRDSA SM2 BIGN ECKCDSA ECSDSA; do
  HARNESS_POOL=512 timeout 240 oracle/_ref/dropin_harness bench libecc_b200/libecc_b200_dropin.so FRP256V1 1048576 $sc 0 2>&1 \
    | grep "DROPIN_BENCH\|HARNESS\|bench rep\|FAIL" >> gpurun_out/r2s3_dropin_ds_bench.log
done
HARNESS_POOL=512 timeout 240 oracle/_ref/dropin_harness bench libecc_b200/libecc_b200_dropin.so FRP256V1 1048576 ECGDSA 64 2>&1 \
    | grep "DROPIN_BENCH\|HARNESS\|bench rep\|FAIL" >> gpurun_out/r2s3_dropin_ds_bench.log
cat gpurun_out/r2s3_dropin_ds_bench.log | cut -c1-330
( time python bench.py --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/r2s3_bench.json 2> gpurun_out/r2s3_bench.err; tail -4 gpurun_out/r2s3_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2s3_bench.json").read().strip().splitlines()[-1])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in l.items() if k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","gpu_launches","parity_spot_check","parity_on_cpu_prefix")})
print("e2e", l["e2e"]["value"], "roofline", l["roofline"]["frac"], l["roofline"]["frac_executed_imad_wide"], "cpu", l["cpu_baseline"]["value"])
print("extra keys", list(l["extra"].keys()))
PY
