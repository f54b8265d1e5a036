"""CPU checks of the ALGORITHMS the kernels run (host build of libecc_b200/csrc/{fp,ec}.cuh, tests/hostsim) against
plain integers, the oracle and the golden vectors.  This is not the product path (that is test_gpu_*.py through
the C ABI); it keeps kernel-algorithm bugs from costing GPU minutes and yields M_impl."""
import numpy as np
import pytest

from common import (hx_fit, ALL_CURVES, CURVES, HASHLEN, ORDER, PRIME, edge_scalars, golden, hostsim_lib, hx, make_signatures,
                    oracle_smul, oracle_verify, random_scalars, rng, _buf)


def be(vals, nbytes):
    return np.frombuffer(b"".join(int(v).to_bytes(nbytes, "big") for v in vals), dtype=np.uint8).copy()


def from_be(arr, nbytes):
    raw = np.ascontiguousarray(arr).tobytes()
    return [int.from_bytes(raw[i:i + nbytes], "big") for i in range(0, len(raw), nbytes)]


def rand_mod(g, mod, n):
    nb = (mod.bit_length() + 7) // 8
    return [int.from_bytes(g.bytes(nb + 8), "big") % mod for _ in range(n)]


@pytest.mark.parametrize("curve", list(ALL_CURVES))
def test_field_ops_against_integers(curve):
    """Pattern of the reference's src/arithmetic_tests (FP_MUL_MONTY / FP_ADD / FP_SUB / FP_INV vs big ints)."""
    lib = hostsim_lib()
    cid, plen, _ = ALL_CURVES[curve]
    g = rng(11)
    rbits = 64 * ((PRIME[curve].bit_length() + 63) // 64)  # the reference's R = 2^(64 * wlen)
    for which, mod in ((0, PRIME[curve]), (1, ORDER[curve])):
        n = 200
        a = rand_mod(g, mod, n) + [0, 1, mod - 1, mod - 1, 0]
        b = rand_mod(g, mod, n) + [0, mod - 1, mod - 1, 1, mod - 1]
        out = np.zeros(len(a) * plen, dtype=np.uint8)
        assert lib.hostsim_fp_mul(cid, which, len(a), _buf(be(a, plen)), _buf(be(b, plen)), _buf(out)) == 0
        rinv = pow(1 << rbits, -1, mod)
        assert from_be(out, plen) == [x * y * rinv % mod for x, y in zip(a, b)]
    p = PRIME[curve]
    a = rand_mod(g, p, 100) + [0, p - 1, 1]
    b = rand_mod(g, p, 100) + [0, p - 1, p - 1]
    out = np.zeros(len(a) * plen, dtype=np.uint8)
    lib.hostsim_fp_op(cid, 0, len(a), _buf(be(a, plen)), _buf(be(b, plen)), _buf(out))
    assert from_be(out, plen) == [(x + y) % p for x, y in zip(a, b)]
    lib.hostsim_fp_op(cid, 1, len(a), _buf(be(a, plen)), _buf(be(b, plen)), _buf(out))
    assert from_be(out, plen) == [(x - y) % p for x, y in zip(a, b)]
    lib.hostsim_fp_op(cid, 2, len(a), _buf(be(a, plen)), _buf(be(b, plen)), _buf(out))
    assert from_be(out, plen) == [pow(x, p - 2, p) for x in a]


@pytest.mark.parametrize("curve", list(ALL_CURVES))
def test_safegcd_inversion_against_integers_and_fermat(curve):
    """Field::inv (Bernstein-Yang division steps on 30-bit limbs) mod p and mod q: against Python's pow(x, -1, m) and
    against the Fermat power it replaced, on random values and on the values that stress the limb / sign handling."""
    lib = hostsim_lib()
    cid, plen, _ = ALL_CURVES[curve]
    g = rng(17)
    for which, mod in ((0, PRIME[curve]), (1, ORDER[curve])):
        bits = mod.bit_length()
        xs = rand_mod(g, mod, 300)
        xs += [0, 1, 2, 3, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, 1 << 29, 1 << 30, (1 << 30) - 1, (1 << 31) + 1,
               1 << 60, (1 << 60) - 1, (1 << (bits - 1)) % mod, ((1 << (bits - 1)) - 1) % mod, (1 << (bits - 2)) + 1]
        xs += [(1 << k) % mod for k in range(29, bits, 61)] + [(mod - (1 << k)) % mod for k in range(1, bits, 53)]
        xs += [pow(3, k, mod) for k in range(1, 40)]
        out = np.zeros(len(xs) * plen, dtype=np.uint8)
        assert lib.hostsim_fp_inv(cid, which, 0, len(xs), _buf(be(xs, plen)), _buf(out)) == 0
        want = [pow(x, -1, mod) if x else 0 for x in xs]
        assert from_be(out, plen) == want
        out2 = np.zeros(64 * plen, dtype=np.uint8)
        assert lib.hostsim_fp_inv(cid, which, 1, 64, _buf(be(xs[:64], plen)), _buf(out2)) == 0
        assert from_be(out2, plen) == want[:64]


@pytest.mark.parametrize("curve", list(ALL_CURVES))
def test_group_law_including_exceptional_cases(curve):
    lib = hostsim_lib()
    cid, plen, qlen = ALL_CURVES[curve]
    q = ORDER[curve]
    ks = [1, 2, 3, 5, q - 1, q - 2, 7]
    pts, st = oracle_smul(curve, be(ks, qlen).reshape(-1, qlen))
    P = {k: pts[i] for i, k in enumerate(ks)}
    zero = np.zeros(2 * plen, dtype=np.uint8)

    def op(which, a, b):
        out = np.zeros(2 * plen, dtype=np.uint8)
        s = np.zeros(1, dtype=np.int8)
        assert lib.hostsim_point_op(cid, which, _buf(a), _buf(b), _buf(out), _buf(s)) == 0
        return out, int(s[0])

    for which in (0, 1, 3):                                                          # full / mixed / XYZZ mixed
        o, s = op(which, P[2], P[3]); assert s == 0 and (o == P[5]).all()            # generic
        o, s = op(which, P[1], P[1]); assert s == 0 and (o == P[2]).all()            # P == Q -> doubling
        o, s = op(which, P[1], P[q - 1]); assert s == 1                              # P == -Q -> infinity
        o, s = op(which, zero, P[3]); assert s == 0 and (o == P[3]).all()            # inf + Q
        o, s = op(which, P[q - 2], P[3]); assert s == 0 and (o == P[1]).all()        # wrap around the order
    o, s = op(0, P[3], zero); assert s == 0 and (o == P[3]).all()                    # P + inf
    o, s = op(0, zero, zero); assert s == 1
    o, s = op(2, P[1], zero); assert s == 0 and (o == P[2]).all()                    # dbl
    o, s = op(2, zero, zero); assert s == 1                                          # dbl(inf) = inf


@pytest.mark.parametrize("curve", list(ALL_CURVES))
@pytest.mark.parametrize("w", [4, 6])
def test_scalar_mult_matches_oracle(curve, w):
    lib = hostsim_lib()
    cid, plen, qlen = ALL_CURVES[curve]
    if w == 6 and plen >= 64:
        pytest.skip("second comb width only for the smaller fields (the CPU table build dominates the suite)")
    sc = np.concatenate([random_scalars(curve, 16, tag=21, below_q=False), edge_scalars(curve)])
    n = sc.shape[0]
    want, wst = oracle_smul(curve, sc)
    out = np.zeros((n, 2 * plen), dtype=np.uint8); st = np.zeros(n, dtype=np.int8)
    assert lib.hostsim_prj_pt_mul_batch(cid, w, n, _buf(sc), None, _buf(out), _buf(st)) == 0
    assert (st == wst).all() and (out == want).all()
    if w == 4:  # variable base once per curve
        good = wst == 0
        pts = want[good].copy()
        sc2 = np.concatenate([edge_scalars(curve), random_scalars(curve, 16, tag=22, below_q=False)])[: pts.shape[0]]
        pts[3, 5] ^= 0x10  # off-curve point -> -1
        want2, wst2 = oracle_smul(curve, sc2, pts)
        out2 = np.zeros_like(want2); st2 = np.zeros_like(wst2)
        assert lib.hostsim_prj_pt_mul_batch(cid, 4, pts.shape[0], _buf(sc2), _buf(pts), _buf(out2), _buf(st2)) == 0
        assert (st2 == wst2).all() and (out2 == want2).all() and wst2[3] == -1


@pytest.mark.parametrize("curve", ["SECP256R1", "SECP384R1", "SECP521R1", "SECP224R1", "SECP192R1"])
def test_ecccdh_kat(curve):
    lib = hostsim_lib()
    cid, plen, qlen = ALL_CURVES[curve]
    vecs = [v for v in golden("ecccdh_kat.json") if v["curve"] == curve]
    d = np.stack([hx_fit(v["priv"], ALL_CURVES[v["curve"]][2]) for v in vecs]); peers = np.stack([hx(v["peer_pub"]) for v in vecs])
    n = len(vecs)
    out = np.zeros((n, 2 * plen), dtype=np.uint8); st = np.zeros(n, dtype=np.int8)
    lib.hostsim_prj_pt_mul_batch(cid, 5, n, _buf(d), None, _buf(out), _buf(st))
    assert [o.tobytes().hex() for o in out] == [v["our_pub"] for v in vecs]
    lib.hostsim_prj_pt_mul_batch(cid, 5, n, _buf(d), _buf(peers), _buf(out), _buf(st))
    assert [o[:plen].tobytes().hex() for o in out] == [v["shared"] for v in vecs]


def test_ecdsa_verify_core_kat_and_wycheproof_sample():
    lib = hostsim_lib()
    for v in golden("ecdsa_kat.json"):
        cid, plen, qlen = ALL_CURVES[v["curve"]]
        out = np.zeros(1, dtype=np.int8)
        lib.hostsim_ecdsa_verify_batch(cid, 4, 1, _buf(hx(v["sig"])), _buf(hx(v["pub"])), _buf(hx(v["digest"])),
                                       HASHLEN[v["hash"]], _buf(out))
        assert out[0] == 0, v["name"]
    for curve in ("SECP256R1", "SECP384R1", "BRAINPOOLP256R1", "SECP256K1", "SECP521R1", "SECP224R1", "BRAINPOOLP512R1"):
        cid, plen, qlen = ALL_CURVES[curve]
        vecs = [v for v in golden("wycheproof_ecdsa.json.gz")
                if v["curve"] == curve and len(v["sig"]) == 4 * qlen and len(v["pub"]) == 4 * plen]
        vecs = vecs[::4]  # a quarter on the CPU; the GPU tests run all of them
        for h in {v["hash"] for v in vecs}:
            vs = [v for v in vecs if v["hash"] == h]
            sig = np.stack([hx(v["sig"]) for v in vs]); pub = np.stack([hx(v["pub"]) for v in vs])
            dg = np.stack([hx(v["digest"]) for v in vs])
            got = np.zeros(len(vs), dtype=np.int8)
            lib.hostsim_ecdsa_verify_batch(cid, 4, len(vs), _buf(sig), _buf(pub), _buf(dg), HASHLEN[h], _buf(got))
            want = np.array([v["ref_verdict"] for v in vs], dtype=np.int8)
            assert (got == want).all(), [v["name"] for v, g_, w_ in zip(vs, got, want) if g_ != w_][:5]


def test_ecdsa_verify_frp256v1_synthetic():
    """FRP256V1 has q > p: exercises the 'candidate r >= p' branch; corrupted tuples must be rejected."""
    lib = hostsim_lib()
    curve = "FRP256V1"
    cid, plen, qlen = ALL_CURVES[curve]
    sigs, pubs, dg, expected = make_signatures(curve, 48, tag=5, corrupt_every=4)
    got = np.zeros(48, dtype=np.int8)
    lib.hostsim_ecdsa_verify_batch(cid, 4, 48, _buf(sigs), _buf(pubs), _buf(dg), 32, _buf(got))
    assert (got == expected).all()
    assert (expected[::4] == -1).all() and (np.delete(expected, np.s_[::4]) == 0).all()


def test_multiplication_counts():
    """M_impl: field multiplications of the implemented algorithms, counted in this host build."""
    lib = hostsim_lib()
    cid, plen, qlen = ALL_CURVES["SECP256R1"]
    sc = random_scalars("SECP256R1", 1, tag=9)
    out = np.zeros(64, dtype=np.uint8); st = np.zeros(1, dtype=np.int8)
    lib.hostsim_prj_pt_mul_batch(cid, 8, 1, _buf(sc), None, _buf(out), _buf(st))
    m_fixed_w8 = lib.hostsim_last_mul_count()
    # 32 windows x 10 (extended-Jacobian mixed add, 8M + 2S; the first add is a copy) + 2 (back to Jacobian) + the
    # normalisation of this host harness (~9 products + the 2 products of the safegcd inversion: no Fermat power)
    assert 31 * 10 + 2 <= m_fixed_w8 <= 32 * 10 + 2 + 20
    pts, _ = oracle_smul("SECP256R1", sc)
    lib.hostsim_prj_pt_mul_batch(cid, 8, 1, _buf(sc), _buf(pts), _buf(out), _buf(st))
    m_var = lib.hostsim_last_mul_count()
    # 64 digits x 4 doublings x 8, a mixed addition (11) for ~15/16 of the digits, the table (6 x 11 + 8 + 5) and its
    # conversion to affine (6 + 12 + 28), and two inversions of this host harness (table + final; safegcd: 2 products
    # each; the kernels share one inversion per 128 threads: roofline.py)
    assert 256 * 8 + 52 * 11 + 125 <= m_var <= 256 * 8 + 64 * 11 + 125 + 30
