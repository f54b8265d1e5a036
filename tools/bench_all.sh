#!/bin/bash
# bench every workload once; prints one compact line each
for wl in secp256r1_fixed_base frp256v1_fixed_base secp384r1_fixed_base secp256r1_variable_base frp256v1_ecdsa_verify secp256r1_ecdsa_verify; do
  python bench.py --workload $wl --steps ${STEPS:-5} --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['value']/1e6,2), 'M/s e2e', round(d['e2e']['value']/1e6,2), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), d['parity_spot_check'])"
done
