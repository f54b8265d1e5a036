#!/bin/bash
# final numbers of the round: N = $NGPU default bench line (what the driver runs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-1}
if [ "$N" = 1 ]; then
  ( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err
else
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/r2_final_n$N.json 2> gpurun_out/r2_final_n$N.err
fi
python - $N <<'PY'
import json,sys
N=sys.argv[1]
f=f"gpurun_out/r2_final_n{N}.json"
try:
    l=json.loads([x for x in open(f).read().strip().splitlines() if x.startswith("{")][-1])
    pr=l.get("per_rank") or {}
    print("N=%s value %.1f M/s ms/step %.4f e2e %.1f K1 %s K4 %s steps %s %s clocks %s bound %s"%(N,l["value"]/1e6,l["ms_per_step"],l["e2e"]["value"]/1e6,
       [round(x,3) for x in pr.get("kernel_ms",[l["roofline"]["kernel_ms"]])],[round(x,3) for x in pr.get("normalisation_ms",[l["roofline"].get("normalisation_kernel_ms") or 0])],
       [round(x,3) for x in pr.get("step_ms",[])],{k:v for k,v in l.items() if k.startswith("gather_") or k.startswith("parity")}, l["clocks"], l.get("host_cpus_bound_near_gpu")))
    print("roofline frac %.3f exec %.3f"%(l["roofline"]["frac"], l["roofline"]["frac_executed_imad_wide"]))
    for k,v in (l.get("extra") or {}).items():
        print("  extra",k,json.dumps({kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","kernel_ms","e2e","e2e_value","parity_spot_check","parity_on_cpu_prefix","seconds_best_of_3","gather_matches_nccl","parity","items_per_gpu","parity_first_items_of_every_shard")})[:420])
    if "cpu_baseline" in l: print("cpu", l["cpu_baseline"]["value"], l["cpu_baseline"]["kind"], l["cpu_baseline"]["cores"])
except Exception as e:
    print("failed",e); print(open(f.replace(".json",".err")).read()[-3000:])
PY
tail -3 gpurun_out/r2_final_n$N.err
