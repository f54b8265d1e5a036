#!/bin/bash
# round 2, validation call (1 GPU): whole GPU suite, launch list + K4 capture, default bench with extras, pipeline trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu9.log 2>&1
tail -8 gpurun_out/r2_pytest_gpu9.log
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -s 320 -c 60 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
grep -c "k_smul_fixed\|k_to_affine" gpurun_out/r02_launches_default.csv
$NCU --set full --import-source on --kernel-name-base demangled -k regex:"k_to_affine<.*\(int\)0>" -s 1 -c 1 -o gpurun_out/r02_k4 -f \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k4.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_k4.ncu-rep > gpurun_out/r02_ncu_k4_to_affine.csv 2> gpurun_out/r02_ncu_k4.err; rm -f gpurun_out/r02_k4.ncu-rep
grep -E "Kernel Name|gpu__time_duration|fmaheavy|registers" gpurun_out/r02_ncu_k4_to_affine.csv | cut -c1-150
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_default9.json 2> gpurun_out/r2_bench_default9.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2_bench_default9.json").read().strip().splitlines()[-1])
print("value %.1f M/s ms/step %.3f k1 %.3f k4 %s e2e %.1f (2^22 %.1f, 2^24 %.1f) traffic %s"%(l["value"]/1e6,l["ms_per_step"],l["roofline"]["kernel_ms"],l["roofline"].get("normalisation_kernel_ms"),l["e2e"]["value"]/1e6,l["extra"]["e2e_2^22"]["value"]/1e6,l["extra"]["e2e_2^24"]["value"]/1e6, (l["roofline"]["traffic"] or {}).get("bytes_per_launch")))
for k,v in l["extra"].items():
    if "value" in v and "kernel_ms" in v: print(k, "%.2f M/s kernel %.3f ms e2e %.2f parity %s cpu %s (%s)"%(v["value"]/1e6,v["kernel_ms"],v["e2e"]["value"]/1e6,v["parity_spot_check"],v.get("parity_on_cpu_prefix"), (v.get("cpu_baseline") or {}).get("value")))
print("cpu", l["cpu_baseline"]["value"], l["cpu_baseline"]["kind"], l["parity_on_cpu_prefix"])
PY
tail -3 gpurun_out/r2_bench_default9.err
ECCB200_PIPE_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2> gpurun_out/r2_pipe_trace9.log; grep "eccb200 pipe" gpurun_out/r2_pipe_trace9.log | tail -8
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
