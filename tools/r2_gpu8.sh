#!/bin/bash
# round 2, 8-GPU call: the scaling point the driver measures at round end, both transports, config 5, multi-device C ABI
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
show() {
python - "$1" "$2" <<'PY'
import json,sys
f,label=sys.argv[1],sys.argv[2]
try:
    l=json.loads([x for x in open(f).read().strip().splitlines() if x.startswith("{")][-1])
    pr=l.get("per_rank") or {}
    print("%s: value %.1f M/s ms/step %.3f e2e %.1f | K1 %s | K4 %s | step %s | %s | clocks %s"%(label,l["value"]/1e6,l["ms_per_step"],l["e2e"]["value"]/1e6,
        [round(x,3) for x in pr.get("kernel_ms",[l["roofline"]["kernel_ms"]])],[round(x,3) for x in pr.get("normalisation_ms",[l["roofline"].get("normalisation_kernel_ms") or 0])],
        [round(x,3) for x in pr.get("step_ms",[])],{k:v for k,v in l.items() if k.startswith("gather_") or k=="parity_spot_check"}, l["clocks"]))
    for k,v in (l.get("extra") or {}).items():
        print("   extra", k, json.dumps(v)[:600])
except Exception as e:
    print(label,"failed",e); print(open(f.replace(".json",".err")).read()[-3000:])
PY
}
N=${NGPU:-8}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/r2_n${N}_default.json 2> gpurun_out/r2_n${N}_default.err
show gpurun_out/r2_n${N}_default.json "N=$N peer-root (default, with extras)"
tail -4 gpurun_out/r2_n${N}_default.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 3 --gather fused-root --no-extra > gpurun_out/r2_n${N}_fused.json 2> gpurun_out/r2_n${N}_fused.err
show gpurun_out/r2_n${N}_fused.json "N=$N fused-root"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 20 --warmup 3 --gather nccl --no-extra > gpurun_out/r2_n${N}_nccl.json 2> gpurun_out/r2_n${N}_nccl.err
show gpurun_out/r2_n${N}_nccl.json "N=$N nccl all_gather per step"
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "all_gpus or peer_gather" ) > gpurun_out/r2_pytest8.log 2>&1
tail -3 gpurun_out/r2_pytest8.log
