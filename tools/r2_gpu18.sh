#!/bin/bash
# session 3, last call: eccb200_prj_pt_unique_batch on its persistent device buffer — harness direct (two curves: the
# buffer grows from 1 item upwards) and the ECGDSA / ECFSDSA adapters at 2^20 with the phase clock
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( HARNESS_CURVES=FRP256V1,SECP384R1 timeout 40 oracle/_ref/dropin_harness direct libecc_b200/libecc_b200_dropin.so 2>&1 | tail -4 ) > gpurun_out/r2s3c.log
for sc in ECGDSA ECFSDSA; do
  ECCB200_DROPIN_TIMING=1 HARNESS_POOL=64 HARNESS_REPS=6 timeout 25 oracle/_ref/dropin_harness bench libecc_b200/libecc_b200_dropin.so FRP256V1 1048576 $sc 0 2>&1 \
    | grep "DROPIN_BENCH\|HARNESS\|bench rep\|FAIL\|timing" >> gpurun_out/r2s3c.log
done
cut -c1-250 gpurun_out/r2s3c.log
