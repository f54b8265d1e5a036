#!/bin/bash
# compute-sanitizer over a small slice of the GPU suite: memcheck (out-of-bounds / misaligned accesses) and racecheck
# (shared-memory hazards: CTA-wide inversion scans, warp-transposed gather stores)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SEL='tests/test_gpu_multi.py::test_fused_gather_on_one_gpu tests/test_gpu_multi.py::test_copy_engine_push_on_one_gpu tests/test_schnorr.py::test_gpu_double_smul_against_oracle tests/test_bip0340.py::test_gpu_bip0340_verify'
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest $SEL -m gpu -x -q -k "SECP256R1 or SECP384R1 or SECP521R1 or push" ) > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/r2_memcheck.log | tail -8
( timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 78 --print-limit 20 python -m pytest tests/test_gpu_multi.py::test_fused_gather_on_one_gpu tests/test_schnorr.py::test_gpu_double_smul_against_oracle -m gpu -x -q -k "SECP256R1" ) > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_racecheck.log | tail -8
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python bench.py --batch-log2 14 --steps 2 --warmup 1 --comb-window 12 --no-extra --no-cpu-baseline ) > gpurun_out/r2_memcheck_bench.log 2>&1
echo "memcheck bench exit $?"; grep -E "ERROR SUMMARY" gpurun_out/r2_memcheck_bench.log | tail -3
