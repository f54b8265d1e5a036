#!/bin/bash
# round 2 evidence captures (1 GPU): launch list of the default bench + ncu --set full of the kernels DESIGN.md cites.
# Numbers printed under ncu are never bench values; only the reports / launch list are used.  The reports are
# summarised on the box (tools/ncu_summary.py) and removed: gpurun_out/ may not exceed 64 MiB.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --clock-control none"
summ() { python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$2.csv 2> gpurun_out/$2.err; [ "$3" = keep ] || rm -f gpurun_out/$1.ncu-rep; wc -l gpurun_out/$2.csv; }
# 1. launch list of the default bench command (shares of the step)
$NCU --metrics gpu__time_duration.sum -c 120 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
# 2. K1 (extended-Jacobian comb, w = 26) and K4 of the device-resident leg
$NCU --set full --import-source on -k regex:k_smul_fixed -s 1 -c 1 -o gpurun_out/r02_k1 -f \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k1.log 2>&1
summ r02_k1 r02_ncu_k1_smul_fixed_w26 keep
$NCU --set full --import-source on -k regex:k_to_affine -s 4 -c 1 -o gpurun_out/r02_k4 -f \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k4.log 2>&1
summ r02_k4 r02_ncu_k4_to_affine
# 3. K3 (config 3) and K2 at 2^18
$NCU --set full --import-source on -k regex:k_ecdsa_verify -s 1 -c 1 -o gpurun_out/r02_k3 -f \
    python bench.py --workload frp256v1_ecdsa_verify --batch-log2 18 --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k3.log 2>&1
summ r02_k3 r02_ncu_k3_verify_frp256v1
$NCU --set full --import-source on -k regex:k_smul_var -s 1 -c 1 -o gpurun_out/r02_k2 -f \
    python bench.py --workload secp256r1_variable_base --batch-log2 18 --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k2.log 2>&1
summ r02_k2 r02_ncu_k2_smul_var
# 4. the roofline denominator and the rejected north-star layout
$NCU --set full -k regex:k_imad_peak -s 2 -c 1 -o gpurun_out/r02_imad_peak -f \
    python -c "import roofline; print(roofline.imad_peak_measured(0))" > gpurun_out/r02_ncu_imad.log 2>&1
summ r02_imad_peak r02_ncu_imad_peak
$NCU --set full -k regex:k_fp_mul_chain -s 2 -c 1 -o gpurun_out/r02_layout_thread -f \
    python tools/microbench_layout.py > gpurun_out/r02_ncu_layout1.log 2>&1
summ r02_layout_thread r02_ncu_layout_thread_per_element
$NCU --set full -k regex:k_fp_mul_striped_chain -s 2 -c 1 -o gpurun_out/r02_layout_striped -f \
    python tools/microbench_layout.py > gpurun_out/r02_ncu_layout2.log 2>&1
summ r02_layout_striped r02_ncu_layout_lane_striped
python tools/microbench_layout.py > gpurun_out/r02_microbench_layout.json 2>&1; cat gpurun_out/r02_microbench_layout.json
du -sh gpurun_out; ls gpurun_out | head -40
