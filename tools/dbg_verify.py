import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, libecc_b200
from common import *
v = golden("ecdsa_kat.json")[0]
for w in (8, 16):
    eng = libecc_b200.Engine(v["curve"], 0, w)
    print("w", w, "verdict code", eng.ecdsa_verify_batch(hx(v["sig"]), hx(v["pub"]), hx(v["digest"]), 32))
# u, v in python
q = ORDER["SECP256R1"]
r = int(v["sig"][:64], 16); s = int(v["sig"][64:], 16); e = int(v["digest"], 16) % q
wv = pow(s, -1, q); u = e * wv % q; vv = r * wv % q
eng = libecc_b200.Engine("SECP256R1", 0, 16)
uG, st = eng.prj_pt_mul_batch(np.frombuffer(u.to_bytes(32, "big"), dtype=np.uint8))
vY, st2 = eng.prj_pt_mul_batch(np.frombuffer(vv.to_bytes(32, "big"), dtype=np.uint8), hx(v["pub"]))
o1, _ = oracle_smul("SECP256R1", np.frombuffer(u.to_bytes(32, "big"), dtype=np.uint8))
o2, _ = oracle_smul("SECP256R1", np.frombuffer(vv.to_bytes(32, "big"), dtype=np.uint8), hx(v["pub"]))
print("uG ok", (uG == o1).all(), "vY ok", (vY == o2).all())
sigs, pubs, dg, expected = make_signatures("SECP256R1", 16, tag=2, corrupt_every=8)
print(eng.ecdsa_verify_batch(sigs, pubs, dg, 32), expected)
