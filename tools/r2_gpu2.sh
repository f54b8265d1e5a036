#!/bin/bash
# round 2, second GPU call (2 GPUs): new tests, drop-in harness, peer gather at N=2 vs N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py "tests/test_gpu_parity.py::test_full_size_ecdsa_verify_frp256v1" -x -q ) > gpurun_out/r2_pytest_multi.log 2>&1
tail -15 gpurun_out/r2_pytest_multi.log
timeout 300 oracle/_ref/dropin_harness threads libecc_b200/libecc_b200_dropin.so > gpurun_out/r2_dropin_threads.log 2>&1; tail -3 gpurun_out/r2_dropin_threads.log
timeout 600 oracle/_ref/dropin_harness bench libecc_b200/libecc_b200_dropin.so FRP256V1 262144 > gpurun_out/r2_dropin_bench.log 2>&1; tail -6 gpurun_out/r2_dropin_bench.log
for g in peer-root peer-all nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --gather $g --no-extra > gpurun_out/r2_scale2_$g.json 2> gpurun_out/r2_scale2_$g.err
  python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/r2_scale2_$g.json").read().strip().splitlines() if x.startswith("{")][-1])
    print("$g N=2 value %.1f M/s ms/step %.3f e2e %.1f k1/rank %s parity %s %s"%(l["value"]/1e6,l["ms_per_step"],l["e2e"]["value"]/1e6,l.get("kernel_ms_per_rank"),l["parity_spot_check"],{k:v for k,v in l.items() if k.startswith("gather_")}))
except Exception as e:
    print("$g failed", e); print(open("gpurun_out/r2_scale2_$g.err").read()[-2500:])
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_scale1.json 2> gpurun_out/r2_scale1.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r2_scale1.json").read().strip().splitlines()[-1])
print("N=1 value %.1f M/s ms/step %.3f e2e %.1f k1 %.3f"%(l["value"]/1e6,l["ms_per_step"],l["e2e"]["value"]/1e6,l["roofline"]["kernel_ms"]))
PY
