#!/bin/bash
# round 2, sixth GPU call (2 GPUs): copy-engine pipelined gather vs fused
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2_pytest6.log 2>&1
tail -4 gpurun_out/r2_pytest6.log
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
    print(label,"failed",e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
}
for g in peer-root peer-all fused-root; do
  ECCB200_GATHER_SLICE_WAVES=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --gather $g --no-extra > gpurun_out/r2_u2_$g.json 2> gpurun_out/r2_u2_$g.err
  show gpurun_out/r2_u2_$g.json "N=2 $g"
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_u1.json 2> gpurun_out/r2_u1.err
show gpurun_out/r2_u1.json "N=1"
