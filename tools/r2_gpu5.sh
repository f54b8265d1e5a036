#!/bin/bash
# round 2, fifth GPU call (2 GPUs): occupancy-sized waves, coalesced gather stores, shaped host pipeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -m gpu -x -q ) > gpurun_out/r2_pytest5.log 2>&1
tail -4 gpurun_out/r2_pytest5.log
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
    for k,v in (l.get("extra") or {}).items():
        if k.startswith("e2e"): print("   ",k,"%.1f M/s"%(v["value"]/1e6), v.get("parity"))
except Exception as e:
    print(label,"failed",e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
}
for waves in 0 4 2; do
  ECCB200_GATHER_SLICE_WAVES=$waves timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --gather peer-root --no-extra > gpurun_out/r2_t2_w$waves.json 2> gpurun_out/r2_t2_w$waves.err
  show gpurun_out/r2_t2_w$waves.json "N=2 peer-root slice_waves=$waves"
done
ECCB200_GATHER_SLICE_WAVES=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --gather peer-all --no-extra > gpurun_out/r2_t2_all.json 2> gpurun_out/r2_t2_all.err
show gpurun_out/r2_t2_all.json "N=2 peer-all unsliced"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t1.json 2> gpurun_out/r2_t1.err
show gpurun_out/r2_t1.json "N=1 plain (shaped host pipeline)"
ECCB200_PIPE_SHAPE=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_t1_eq.json 2> gpurun_out/r2_t1_eq.err
show gpurun_out/r2_t1_eq.json "N=1 plain (equal chunks)"
ECCB200_PIPE_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2> gpurun_out/r2_pipe_trace.log; grep "eccb200 pipe" gpurun_out/r2_pipe_trace.log | tail -12
