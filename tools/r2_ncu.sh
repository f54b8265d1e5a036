#!/bin/bash
# round 2 evidence captures (1 GPU): launch list of the default bench + ncu --set full of the kernels DESIGN.md cites.
# Numbers printed under ncu are never bench values; only the .ncu-rep / launch list are used.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of the default bench command (shares of the step)
$NCU --metrics gpu__time_duration.sum -c 120 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
# 2. K1 (extended-Jacobian comb, w = 26) and K4 of the device-resident leg
$NCU --set full --import-source on -k regex:k_smul_fixed -s 1 -c 1 -o gpurun_out/r02_k1 -f \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k1.log 2>&1
$NCU --set full --import-source on -k regex:k_to_affine -s 4 -c 1 -o gpurun_out/r02_k4 -f \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k4.log 2>&1
# 3. K3 (config 3) and K2 at 2^18
$NCU --set full --import-source on -k regex:k_ecdsa_verify -s 1 -c 1 -o gpurun_out/r02_k3 -f \
    python bench.py --workload frp256v1_ecdsa_verify --batch-log2 18 --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k3.log 2>&1
$NCU --set full --import-source on -k regex:k_smul_var -s 1 -c 1 -o gpurun_out/r02_k2 -f \
    python bench.py --workload secp256r1_variable_base --batch-log2 18 --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k2.log 2>&1
# 4. the roofline denominator and the rejected north-star layout
$NCU --set full -k regex:k_imad_peak -s 2 -c 1 -o gpurun_out/r02_imad_peak -f \
    python -c "import roofline; print(roofline.imad_peak_measured(0))" > gpurun_out/r02_ncu_imad.log 2>&1
$NCU --set full -k regex:k_fp_mul -c 12 -o gpurun_out/r02_layout -f \
    python tools/microbench_layout.py > gpurun_out/r02_ncu_layout.log 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/microbench_layout.py > gpurun_out/r02_microbench_layout.json 2>&1; cat gpurun_out/r02_microbench_layout.json
