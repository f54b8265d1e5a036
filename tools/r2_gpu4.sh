#!/bin/bash
# round 2, fourth GPU call (2 GPUs): sliced gather vs unsliced, per-rank timings; BIP0340 GPU tests; N=1 self-gather overhead
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_bip0340.py tests/test_gpu_multi.py tests/test_gpu_dropin.py -m gpu -x -q ) > gpurun_out/r2_pytest4.log 2>&1
tail -4 gpurun_out/r2_pytest4.log
show() {
python - "$1" "$2" <<'PY'
import json,sys
f,label=sys.argv[1],sys.argv[2]
try:
    l=json.loads([x for x in open(f).read().strip().splitlines() if x.startswith("{")][-1])
    pr=l.get("per_rank") or {}
    print("%s: value %.1f M/s ms/step %.3f e2e %.1f | K1 %s | K4 %s | step %s | %s"%(label,l["value"]/1e6,l["ms_per_step"],l["e2e"]["value"]/1e6,
        [round(x,3) for x in pr.get("kernel_ms",[l["roofline"]["kernel_ms"]])],[round(x,3) for x in pr.get("normalisation_ms",[l["roofline"].get("normalisation_kernel_ms") or 0])],
        [round(x,3) for x in pr.get("step_ms",[])],{k:v for k,v in l.items() if k.startswith("gather_") or k=="parity_spot_check"}))
except Exception as e:
    print(label,"failed",e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
}
for waves in 4 2 0; do
  ECCB200_GATHER_SLICE_WAVES=$waves timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --gather peer-root --no-extra > gpurun_out/r2_s2_w$waves.json 2> gpurun_out/r2_s2_w$waves.err
  show gpurun_out/r2_s2_w$waves.json "N=2 peer-root slice_waves=$waves"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_s1.json 2> gpurun_out/r2_s1.err
show gpurun_out/r2_s1.json "N=1 plain"
for waves in 4 2; do
  BENCH_FORCE_GATHER=1 ECCB200_GATHER_SLICE_WAVES=$waves timeout 600 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --gather peer-root > gpurun_out/r2_s1_self$waves.json 2> gpurun_out/r2_s1_self$waves.err
  show gpurun_out/r2_s1_self$waves.json "N=1 self-gather slice_waves=$waves"
done
