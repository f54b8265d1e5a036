"""One device-resident ECFSDSA multi-scalar-multiplication batch verification at 2^20 (for the ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hashlib
import numpy as np
import torch
import libecc_b200
from common import ALL_CURVES, ORDER

curve = "FRP256V1"
_, plen, qlen = ALL_CURVES[curve]
q = ORDER[curve]
n, m = 1 << int(os.environ.get("MSM_LOG2", "20")), 1 << 12
g = np.random.default_rng(5)
eng = libecc_b200.Engine(curve, comb_window=12)
d = g.integers(0, 256, (m, qlen), dtype=np.uint8); d[:, 0] &= 0x7F; d[:, -1] |= 1
k = g.integers(0, 256, (m, qlen), dtype=np.uint8); k[:, 0] &= 0x7F; k[:, -1] |= 1
pubs, _ = eng.prj_pt_mul_batch(d)
W, _ = eng.prj_pt_mul_batch(k)
sigs = np.zeros((m, 2 * plen + qlen), np.uint8); dg = np.zeros((m, 32), np.uint8)
sigs[:, :2 * plen] = W
for i in range(m):
    h = hashlib.sha256(W[i].tobytes() + b"x").digest()
    dg[i] = np.frombuffer(h, np.uint8)
    s_i = (int.from_bytes(k[i].tobytes(), "big") + int.from_bytes(h, "big") * int.from_bytes(d[i].tobytes(), "big")) % q
    sigs[i, 2 * plen:] = np.frombuffer(s_i.to_bytes(qlen, "big"), np.uint8)
r = n // m
dS, dP, dD = (torch.from_numpy(np.tile(a, (r, 1))).cuda() for a in (sigs, pubs, dg))
for _ in range(int(os.environ.get("MSM_REPS", "2"))):
    ok = eng.ecfsdsa_verify_msm_batch_dev(n, dS.data_ptr(), dP.data_ptr(), dD.data_ptr(), 32, None, 0)
print("verdict", ok)
