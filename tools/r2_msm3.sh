#!/bin/bash
# drop-in ECFSDSA / BIP0340 verify_batch adapters with the K6 fast path: tests + throughput on 2^20 real structs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/r2_msm3_tests.log
H=oracle/_ref/dropin_harness; D=libecc_b200/libecc_b200_dropin.so
( timeout 600 $H bench $D FRP256V1 1048576 ECFSDSA 0 | tail -5
  timeout 600 $H bench $D FRP256V1 1048576 ECFSDSA 64 | tail -2
  ECCB200_DROPIN_MSM_MIN=0 timeout 600 $H bench $D FRP256V1 1048576 ECFSDSA 0 | tail -2
  timeout 600 $H bench $D SECP256K1 1048576 BIP0340 0 | tail -2
  timeout 600 $H bench $D SECP256K1 1048576 BIP0340 64 | tail -2 ) > gpurun_out/r2_msm3_bench.log 2>&1
cat gpurun_out/r2_msm3_tests.log gpurun_out/r2_msm3_bench.log
