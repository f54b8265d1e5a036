#!/bin/bash
# e2e pipeline variants (1 GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {
  label="$1"; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_e2e_$label.json 2> gpurun_out/r2_e2e_$label.err
  python - "$label" <<'PY'
import json,sys
lab=sys.argv[1]
try:
    l=json.loads(open(f"gpurun_out/r2_e2e_{lab}.json").read().strip().splitlines()[-1])
    print("%s: value %.1f e2e 2^20 %.1f 2^22 %.1f 2^24 %.1f | p384 e2e %.1f"%(lab,l["value"]/1e6,l["e2e"]["value"]/1e6,l["extra"]["e2e_2^22"]["value"]/1e6,l["extra"]["e2e_2^24"]["value"]/1e6,l["extra"]["secp384r1_fixed_base"]["e2e"]["value"]/1e6))
except Exception as e:
    print(lab,"failed",e); print(open(f"gpurun_out/r2_e2e_{lab}.err").read()[-1500:])
PY
}
run default X=1
run k4inline ECCB200_PIPE_K4_INLINE=1
run k4inline_eq ECCB200_PIPE_K4_INLINE=1 ECCB200_PIPE_SHAPE=0
run eq ECCB200_PIPE_SHAPE=0
run k4inline_aff2 ECCB200_PIPE_K4_INLINE=1 ECCB200_AFFINE_CTAS=2
