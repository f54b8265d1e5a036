#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m | head -12 > gpurun_out/r2_topo_b.txt
run() {
  label=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 --no-extra > gpurun_out/r2_f8_$label.json 2> gpurun_out/r2_f8_$label.err
  python - $label <<'PY'
import json,sys
lab=sys.argv[1]
f=f"gpurun_out/r2_f8_{lab}.json"
try:
    l=json.loads([x for x in open(f).read().strip().splitlines() if x.startswith("{")][-1])
    pr=l.get("per_rank") or {}
    print("%s: value %.1f M/s ms/step %.4f e2e %.1f K4 %s steps %s p2p %s nccl %s"%(lab,l["value"]/1e6,l["ms_per_step"],l["e2e"]["value"]/1e6,
       [round(x,3) for x in pr.get("normalisation_ms",[])],[round(x,3) for x in pr.get("step_ms",[])], l.get("p2p_push_GBps_per_rank_alone"), l.get("gather_matches_nccl")))
except Exception as e:
    print(lab,"failed",e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
}
run probe BENCH_P2P_PROBE=1
run again X=1
