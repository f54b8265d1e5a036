#!/bin/bash
# session 3, second call: the double-scalar adapters with page-locked key staging; phase timing; drop-in GPU tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu ) > gpurun_out/r2s3b_pytest.log 2>&1; tail -5 gpurun_out/r2s3b_pytest.log
: > gpurun_out/r2s3b_dropin_ds_bench.log
for sc in ECGDSA ECRDSA SM2 BIGN ECKCDSA ECSDSA; do
  ECCB200_DROPIN_TIMING=1 HARNESS_POOL=256 timeout 200 oracle/_ref/dropin_harness bench libecc_b200/libecc_b200_dropin.so FRP256V1 1048576 $sc 0 2>&1 \
    | grep "DROPIN_BENCH\|HARNESS\|bench rep\|FAIL\|timing" >> gpurun_out/r2s3b_dropin_ds_bench.log
done
for sc in SM2 ECGDSA; do
  echo "threads=64" >> gpurun_out/r2s3b_dropin_ds_bench.log
  ECCB200_DROPIN_THREADS=64 ECCB200_DROPIN_TIMING=1 HARNESS_POOL=256 timeout 200 oracle/_ref/dropin_harness bench libecc_b200/libecc_b200_dropin.so FRP256V1 1048576 $sc 0 2>&1 \
    | grep "DROPIN_BENCH\|HARNESS\|bench rep\|FAIL\|timing" >> gpurun_out/r2s3b_dropin_ds_bench.log
done
cut -c1-260 gpurun_out/r2s3b_dropin_ds_bench.log
