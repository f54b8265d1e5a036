#!/usr/bin/env python3
"""DESIGN.md §3 evidence: Montgomery products per second with one thread per element (production) versus the words
of an element striped across 8 lanes with __shfl_sync carries.  Run on a B200: python tools/microbench_layout.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import libecc_b200  # noqa: E402

out = {}
for curve in ("SECP256R1", "FRP256V1"):
    eng = libecc_b200.Engine(curve, 0, 8)
    n, iters = 1 << 20, 256
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] &= 0x7F
    b = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); b[:, 0] &= 0x7F
    r0, ms0 = eng.fp_mul_chain_bench(a, b, iters, striped=False)
    r1, ms1 = eng.fp_mul_chain_bench(a, b, iters, striped=True)
    assert (r0 == r1).all()
    out[curve] = {"thread_per_element_Gmul_s": n * iters / ms0 / 1e6, "lane_striped_Gmul_s": n * iters / ms1 / 1e6,
                  "ratio": ms1 / ms0, "n": n, "iters": iters}
    eng.close()
print(json.dumps(out))
