#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in $LIBS; do
  for wl in secp256r1_fixed_base frp256v1_fixed_base secp384r1_fixed_base secp521r1_fixed_base; do
    ECCB200_LIB=$PWD/libecc_b200/libecc_b200$lib.so timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_ool${lib}_$wl.json 2> gpurun_out/r2_ool${lib}_$wl.err
    python - "$lib" "$wl" <<'PY'
import json,sys
lib,wl=sys.argv[1],sys.argv[2]
try:
    l=json.loads(open(f"gpurun_out/r2_ool{lib}_{wl}.json").read().strip().splitlines()[-1])
    print("lib%s %s: value %.2f M/s K1 %.4f ms parity %s"%(lib or "(inline)",wl,l["value"]/1e6,l["roofline"]["kernel_ms"],l["parity_spot_check"]))
except Exception as e:
    print(lib,wl,"failed",e); print(open(f"gpurun_out/r2_ool{lib}_{wl}.err").read()[-800:])
PY
  done
done
