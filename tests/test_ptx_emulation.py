"""Executes the generated inline-PTX instruction streams of libecc_b200/csrc/fp_ptx.cuh (tools/gen_fp_ptx.py) in the
generator's PTX interpreter and compares with Python integers: validates the carry-chain construction of the
IMAD.WIDE Montgomery multiplier (no dropped or rippling carries: the interpreter asserts on any carry lost by a
non-.cc instruction) without a GPU.  The GPU test test_fp_mul_monty_against_integers then checks the real thing."""
import os
import random
import sys

import pytest

from common import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_fp_ptx  # noqa: E402


def env(n, a, b):
    e = {}
    for k in range(n):
        e[f"a{k}"] = (a >> (32 * k)) & 0xFFFFFFFF
        e[f"b{k}"] = (b >> (32 * k)) & 0xFFFFFFFF
    return e


def result(n, r):
    return sum(r[f"r{k}"] << (32 * k) for k in range(n))


@pytest.mark.parametrize("tag,n,mod", list(gen_fp_ptx.fields()))
def test_generated_streams(tag, n, mod):
    rnd = random.Random(hash(tag) & 0xFFFF)
    add, sub = gen_fp_ptx.gen_add(n, mod), gen_fp_ptx.gen_sub(n, mod)
    # every multiplier variant the header carries (generic, and the add-form reduction of 0xffffffff modulus words for
    # the inlined / out-of-line multipliers), plus the all-words form whatever the configured shares are
    progs = [(gen_fp_ptx.gen_mul(n, mod), gen_fp_ptx.gen_sqr(n, mod))]
    progs += [(m, s) for _, m, s in gen_fp_ptx.variants(n, mod)]
    full = frozenset(gen_fp_ptx.solinas_words(mod, n, n))
    progs.append((gen_fp_ptx.gen_mul(n, mod, full), gen_fp_ptx.gen_sqr(n, mod, full)))
    mul, sqr = progs[0]
    rinv = pow(1 << (32 * n), -1, mod)
    special = [0, 1, 2, mod - 1, mod - 2, (1 << (32 * n - 1)) % mod, (mod >> 1), (1 << 32) - 1, ((1 << 32) - 1) << 32,
               mod - (1 << 32), int("ffffffff" * n, 16) % mod, int("ffffffff" * n, 16) - mod if int("ffffffff" * n, 16) - mod < mod else 5]
    pairs = [(x, y) for x in special for y in special]
    pairs += [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(600)]
    for a, b in pairs:
        for pm, ps in progs:
            assert result(n, pm.run(env(n, a, b))) == a * b * rinv % mod
            assert result(n, ps.run(env(n, a, b))) == a * a * rinv % mod
        assert result(n, add.run(env(n, a, b))) == (a + b) % mod
        assert result(n, sub.run(env(n, a, b))) == (a - b) % mod
    wide, total = mul.count()
    m0_is_one = (-pow(mod, -1, 1 << 32)) % (1 << 32) == 1
    assert wide <= 2 * n * n and total <= 2 * (2 * n * n) + 8 * n + (0 if m0_is_one else n) + 8
    swide, _ = sqr.count()
    assert swide <= n * (n + 1) // 2 + n * n   # symmetric product + full reduction


def test_header_is_up_to_date():
    """fp_ptx.cuh on disk is what the generator produces now."""
    path = os.path.join(ROOT, "libecc_b200", "csrc", "fp_ptx.cuh")
    before, st = open(path).read(), os.stat(path)
    gen_fp_ptx.main()
    assert open(path).read() == before
    os.utime(path, ns=(st.st_atime_ns, st.st_mtime_ns))   # same bytes: keep the timestamp, or every test run forces a rebuild
