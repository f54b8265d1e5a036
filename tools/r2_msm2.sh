#!/bin/bash
# K6 experiment: multiplier inlined in the MSM translation unit vs the default (out-of-line calls); ncu --set full of the accumulation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run_leg() { timeout 600 python -c "
import json, bench
r = bench.run_schnorr_msm(0, 'ecfsdsa', with_cpu=False)
print(json.dumps({k: r[k] for k in ('value', 'ms_per_step', 'accepts_valid_batch', 'rejects_one_flipped_bit')}))
" 2>&1 | tail -2; }
echo "default:"; run_leg
echo "inline:"; ECCB200_LIB=$PWD/libecc_b200/libecc_b200_inl.so run_leg
ECCB200_LIB=$PWD/libecc_b200/libecc_b200_inl.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"k_msm" --csv \
  --log-file gpurun_out/r2_msm_launches_inl.csv python tools/msm_once.py > gpurun_out/r2_msm_ncu_inl.log 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_msm_launches_inl.csv")) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:][-14:]:
    print(r[ki][:60], r[vi])
P
MSM_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_msm_accumulate" -c 1 \
  -o gpurun_out/r2_msm_acc python tools/msm_once.py > gpurun_out/r2_msm_ncu_full.log 2>&1
ncu -i gpurun_out/r2_msm_acc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_msm_accumulate.csv 2>/dev/null
rm -f gpurun_out/r2_msm_acc.ncu-rep
python - <<'P'
import csv
rows = list(csv.reader(open("gpurun_out/r2_ncu_msm_accumulate.csv")))
hdr, val = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__average_warp_latency_issue_stalled_no_instruction.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w: print(w, val[i])
P
