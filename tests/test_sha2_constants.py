"""tools/gen_sha2_constants.py derives the FIPS 180-4 constants from their definition; here a straightforward Python
SHA-256 / SHA-512 / SHA-384 built on those values must agree with hashlib, and the committed .inc must be current."""
import hashlib
import os
import sys

from common import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_sha2_constants as G  # noqa: E402


def _sha2(msg, k, iv, wbits, rounds, rot, outlen):
    mask = (1 << wbits) - 1
    rotr = lambda x, n: ((x >> n) | (x << (wbits - n))) & mask
    blk = wbits * 2
    ml = len(msg) * 8
    msg = msg + b"\x80"
    msg += b"\x00" * ((-len(msg) - blk // 8) % blk)
    msg += ml.to_bytes(blk // 8, "big")
    h = list(iv)
    wb = wbits // 8
    for off in range(0, len(msg), blk):
        w = [int.from_bytes(msg[off + wb * i: off + wb * (i + 1)], "big") for i in range(16)]
        for t in range(16, rounds):
            s0 = rotr(w[t - 15], rot[0]) ^ rotr(w[t - 15], rot[1]) ^ (w[t - 15] >> rot[2])
            s1 = rotr(w[t - 2], rot[3]) ^ rotr(w[t - 2], rot[4]) ^ (w[t - 2] >> rot[5])
            w.append((w[t - 16] + s0 + w[t - 7] + s1) & mask)
        a, b, c, d, e, f, g, hh = h
        for t in range(rounds):
            S1 = rotr(e, rot[6]) ^ rotr(e, rot[7]) ^ rotr(e, rot[8])
            ch = (e & f) ^ (~e & mask & g)
            t1 = (hh + S1 + ch + k[t] + w[t]) & mask
            S0 = rotr(a, rot[9]) ^ rotr(a, rot[10]) ^ rotr(a, rot[11])
            t2 = (S0 + ((a & b) ^ (a & c) ^ (b & c))) & mask
            hh, g, f, e, d, c, b, a = g, f, e, (d + t1) & mask, c, b, a, (t1 + t2) & mask
        h = [(x + y) & mask for x, y in zip(h, (a, b, c, d, e, f, g, hh))]
    return b"".join(x.to_bytes(wb, "big") for x in h)[:outlen]


def test_constants_reproduce_hashlib_and_inc_is_current():
    path = os.path.join(ROOT, "libecc_b200", "csrc", "sha2_constants.inc")
    before, st = open(path).read(), os.stat(path)
    k256, h256, k512, h512, h384 = G.main()
    assert open(path).read() == before
    os.utime(path, ns=(st.st_atime_ns, st.st_mtime_ns))   # same bytes: keep the timestamp (no needless rebuild)
    r256 = (7, 18, 3, 17, 19, 10, 6, 11, 25, 2, 13, 22)
    r512 = (1, 8, 7, 19, 61, 6, 14, 18, 41, 28, 34, 39)
    for m in (b"", b"abc", b"a" * 55, b"a" * 56, b"b" * 64, b"c" * 111, b"d" * 112, b"e" * 300):
        assert _sha2(m, k256, h256, 32, 64, r256, 32) == hashlib.sha256(m).digest()
        assert _sha2(m, k512, h512, 64, 80, r512, 64) == hashlib.sha512(m).digest()
        assert _sha2(m, k512, h384, 64, 80, r512, 48) == hashlib.sha384(m).digest()
