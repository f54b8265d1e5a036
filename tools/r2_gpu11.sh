#!/bin/bash
# register-cap sweep of K2 / K3 after the safegcd change (1 GPU) + drop-in harness with ECSDSA / ECOSDSA / ECKCDSA + K4 capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_schnorr.py tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2_pytest11.log 2>&1
tail -6 gpurun_out/r2_pytest11.log
for lib in "" _mv5 _mv7 _mv8; do
  for wl in frp256v1_ecdsa_verify secp256r1_variable_base; do
    ECCB200_LIB=$PWD/libecc_b200/libecc_b200$lib.so timeout 600 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline --no-extra > gpurun_out/r2_sweep${lib}_$wl.json 2> gpurun_out/r2_sweep${lib}_$wl.err
    python - "$lib" "$wl" <<'PY'
import json,sys
lib,wl=sys.argv[1],sys.argv[2]
try:
    l=json.loads(open(f"gpurun_out/r2_sweep{lib}_{wl}.json").read().strip().splitlines()[-1])
    print("lib%s %s: value %.3f M/s kernel %.3f ms parity %s"%(lib or "(default 6/7)",wl,l["value"]/1e6,l["roofline"]["kernel_ms"],l["parity_spot_check"]))
except Exception as e:
    print(lib,wl,"failed",e); print(open(f"gpurun_out/r2_sweep{lib}_{wl}.err").read()[-800:])
PY
  done
done
NCU="ncu --clock-control none"
$NCU --set full --import-source on --kernel-name-base demangled -k regex:"k_to_affine<.*\(int\)0>" -s 1 -c 1 -o gpurun_out/r02_k4 -f \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_ncu_k4.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_k4.ncu-rep > gpurun_out/r02_ncu_k4_to_affine.csv 2> gpurun_out/r02_ncu_k4.err; rm -f gpurun_out/r02_k4.ncu-rep
grep -E "Kernel Name|gpu__time_duration|fmaheavy|registers" gpurun_out/r02_ncu_k4_to_affine.csv | cut -c1-150
