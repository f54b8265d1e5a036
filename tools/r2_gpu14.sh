#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for a in 4 5 6 8; do
  ECCB200_AFFINE_CTAS=$a timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra 2>gpurun_out/r2_aff$a.err | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AFFINE_CTAS=$a value %.1f ms/step %.4f K1 %.4f K4 %.4f e2e %.1f'%(l['value']/1e6,l['ms_per_step'],l['roofline']['kernel_ms'],l['roofline']['normalisation_kernel_ms'],l['e2e']['value']/1e6))"
done
