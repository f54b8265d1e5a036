"""The drop-in's host arithmetic modulo the group order (struct ModQ in libecc_b200/csrc/dropin.cpp: the nn_mod_mul /
nn_modinv / nn_mod_add / nn_mod of the ECGDSA, ECRDSA, SM2 and BIGN scalar preparations, src/sig/ecgdsa.c:545-569,
src/sig/ecrdsa.c:550-570, src/sig/sm2.c:657-684, src/sig/bign_common.c:898-916) against Python integers on all eleven
curves: 4, 6, 8 and 9 limbs of 64 bits, orders that fill their top limb and orders that do not."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from common import ALL_CURVES, ORDER, ROOT, engine_stub_so, rng, _buf

HOOK_SO = os.path.join(ROOT, "tests", "hostsim", "_build", "libmodq_hook.so")


def hook():
    src = os.path.join(ROOT, "tests", "hostsim", "modq_hook.cpp")
    deps = [src, os.path.join(ROOT, "libecc_b200", "csrc", "dropin.cpp"), os.path.join(ROOT, "libecc_b200", "csrc", "fp.cuh")]
    stub = engine_stub_so()
    if not os.path.exists(HOOK_SO) or os.path.getmtime(HOOK_SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-x", "c++", src, "-o", HOOK_SO,
                        "-L" + os.path.dirname(stub), "-lecc_b200_stub", "-Wl,-rpath," + os.path.dirname(stub), "-ldl",
                        "-lpthread"], check=True, capture_output=True)
    return ctypes.CDLL(HOOK_SO)


def be(v, n):
    return np.frombuffer(int(v).to_bytes(n, "big"), dtype=np.uint8).copy()


def val(arr):
    return int.from_bytes(np.ascontiguousarray(arr).tobytes(), "big")


@pytest.mark.parametrize("curve", list(ALL_CURVES))
def test_mod_q_operations_against_integers(curve):
    lib = hook()
    cid, _, qlen = ALL_CURVES[curve]
    q = ORDER[curve]
    assert lib.modq_order_bits(cid) == q.bit_length()
    g = rng(41)
    rbits = 64 * ((q.bit_length() + 63) // 64)
    vals = [0, 1, 2, q - 1, q - 2, (1 << (q.bit_length() - 1)), (1 << 64) - 1, 1 << 64] + [
        int.from_bytes(g.bytes(qlen + 8), "big") % q for _ in range(60)]
    out = np.zeros(2 * qlen, dtype=np.uint8)
    for i, a in enumerate(vals):
        b = vals[(7 * i + 3) % len(vals)]
        for op, want in ((0, (a + b) % q), (1, (-a) % q), (2, a * b % q), (4, (a << rbits) % q),
                         (5, a * pow(1 << rbits, -1, q) % q)):
            assert lib.modq_op(cid, op, _buf(be(a, qlen)), _buf(be(b, qlen)), _buf(out)) == 0
            assert val(out[:qlen]) == want, (curve, op, hex(a), hex(b))
        if a and b:
            assert lib.modq_op(cid, 3, _buf(be(a, qlen)), _buf(be(b, qlen)), _buf(out)) == 0
            assert val(out[:qlen]) == pow(a, -1, q) and val(out[qlen:]) == pow(b, -1, q)


@pytest.mark.parametrize("curve", list(ALL_CURVES))
def test_simultaneous_inversion(curve):
    """Montgomery's trick over a chunk of items, as the ECGDSA (r^-1) and ECRDSA (h^-1) adapters use it."""
    lib = hook()
    cid, _, qlen = ALL_CURVES[curve]
    q = ORDER[curve]
    g = rng(43)
    for k in (1, 2, 5, 257):
        xs = [1, q - 1][:min(k, 2)] + [int.from_bytes(g.bytes(qlen + 8), "big") % (q - 1) + 1 for _ in range(max(0, k - 2))]
        buf = np.concatenate([be(x, qlen) for x in xs])
        out = np.zeros(k * qlen, dtype=np.uint8)
        assert lib.modq_inv_many(cid, _buf(buf), k, _buf(out)) == 0
        assert [val(out[i * qlen:(i + 1) * qlen]) for i in range(k)] == [pow(x, -1, q) for x in xs]


@pytest.mark.parametrize("curve", list(ALL_CURVES))
def test_reduction_of_byte_strings_and_digest_truncation(curve):
    lib = hook()
    cid, _, qlen = ALL_CURVES[curve]
    q = ORDER[curve]
    g = rng(47)
    out = np.zeros(qlen, dtype=np.uint8)
    for ln in (0, 1, 7, 8, 9, qlen - 1, qlen, qlen + 1, 8 * ((q.bit_length() + 63) // 64), 8 * ((q.bit_length() + 63) // 64) + 1,
               48, 64, 65, 128):
        for raw in (g.bytes(ln), b"\xff" * ln, b"\x00" * ln):
            arr = np.frombuffer(raw, dtype=np.uint8).copy() if ln else np.zeros(1, dtype=np.uint8)
            assert lib.modq_reduce(cid, _buf(arr), ln, _buf(out)) == 0
            assert val(out) == int.from_bytes(raw, "big") % q, (curve, ln)
    for hlen in (20, 28, 32, 48, 64, 66, 128):
        for raw in (g.bytes(hlen), b"\xff" * hlen, b"\x80" + b"\x00" * (hlen - 1)):
            arr = np.frombuffer(raw, dtype=np.uint8).copy()
            assert lib.modq_digest_truncated(cid, _buf(arr), hlen, _buf(out)) == 0
            shift = max(0, 8 * hlen - q.bit_length())  # src/sig/ecgdsa.c:545-556
            assert val(out) == (int.from_bytes(raw, "big") >> shift) % q, (curve, hlen)
