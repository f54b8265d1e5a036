"""ECFSDSA verification rate on one GPU: 2^20 signatures (256 made by the unmodified reference, tiled,
every 4th corrupted), device-resident buffers, CUDA events around eccb200_ecfsdsa_verify_batch_dev."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libecc_b200  # noqa: E402
from test_ecfsdsa import workload  # noqa: E402

curve = sys.argv[1] if len(sys.argv) > 1 else "FRP256V1"
n = 1 << 20
sigs, pubs, dg, hlen, want = workload(curve, 256, 7500)
reps = n // 256
eng = libecc_b200.Engine(curve)
dev = torch.device("cuda:0")
d_s = torch.from_numpy(np.tile(sigs, (reps, 1))).to(dev)
d_p = torch.from_numpy(np.tile(pubs, (reps, 1))).to(dev)
d_d = torch.from_numpy(np.tile(dg, (reps, 1))).to(dev)
d_v = torch.zeros(n, dtype=torch.int8, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def call():
    rc = eng.lib.eccb200_ecfsdsa_verify_batch_dev(eng._h, n, d_s.data_ptr(), d_p.data_ptr(), d_d.data_ptr(), hlen,
                                                  d_v.data_ptr(), stream)
    assert rc == 0


for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
assert (d_v.cpu().numpy() == np.tile(want, reps)).all()
print(f"{curve} ECFSDSA verify: {n / ms / 1e3:.2f} M/s ({ms:.2f} ms per 2^20, verdicts match the reference)")
