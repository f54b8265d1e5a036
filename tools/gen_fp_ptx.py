#!/usr/bin/env python3
"""Generates libecc_b200/csrc/fp_ptx.cuh: fully unrolled inline-PTX Montgomery multiplication / addition /
subtraction for every field of curve_constants.inc, and carries a small PTX interpreter so that the generated
instruction streams can be executed on the CPU against Python integers (tests/test_ptx_emulation.py).

Algorithm of `mul` (the device restatement of nn_mul_redc1, /root/reference/src/nn/nn_mul_redc1.c:124-218):
word-serial interleaved Montgomery multiplication (CIOS) on N 32-bit words, arranged so that every 32x32->64
product is ONE IMAD.WIDE with carry-in/out:

  * the running sum lives in two register files E and O indexed by word position; E is only ever used as the 64-bit
    pairs (k, k+1) with k even, O as the pairs with k odd, so each register keeps one pair alignment for its whole
    life and `mad.lo.cc / madc.hi.cc` pairs fuse into IMAD.WIDE.U32(.X) on aligned register pairs;
  * row i (multiplier word b_i, then Montgomery quotient m_i) adds a_j*b_i and p_j*m_i at position i+j: the even
    j's form one carry chain in X = (i even ? E : O), the odd j's one chain in Y = the other file;
  * nothing is shifted: row i simply starts one position higher (register renaming is free when fully unrolled);
  * before m_i is taken, the leftover word of the other file at position i is folded in (add.cc) and its carry
    feeds the chain that starts at position i+1;
  * every chain's carry-out lands in a register that so far holds only carries (bounded by 3), so no carry ever
    ripples and none is dropped — this is what makes the scheme valid for full-width moduli (2^(32N-1) < p).

Work per multiplication: 2N^2 IMAD.WIDE (+ N for m_i when M0 != 1) and ~7N carry/merge/select instructions on
the ALU pipe.  Words of the modulus that are 0 or 1 are specialised away (P-256: 4 of 8, P-384: 2 of 12).
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_curve_constants import CURVES, nwords  # noqa: E402

MASK = 0xFFFFFFFF


class Prog:
    """A straight-line PTX program over 32-bit registers and the carry flag."""

    def __init__(self):
        self.ins = []      # (op, dst, [srcs]) ; srcs are register names or ints
        self.regs = set()

    def emit(self, op, dst, *srcs):
        self.ins.append((op, dst, list(srcs)))
        if dst is not None:
            self.regs.add(dst)
        for s in srcs:
            if isinstance(s, str):
                self.regs.add(s)

    # ---- interpreter ---------------------------------------------------------------------------------------
    def run(self, env):
        """Execute on a dict of register -> int (inputs preset). Returns the dict."""
        r = dict(env)
        cc = 0
        pred = {}

        def val(s):
            return s if isinstance(s, int) else r[s]

        for op, dst, srcs in self.ins:
            v = [val(s) for s in srcs if s != "pq"]
            if op == "mul.lo.u32":
                r[dst] = (v[0] * v[1]) & MASK
            elif op == "mul.hi.u32":
                r[dst] = (v[0] * v[1]) >> 32
            elif op in ("mad.lo.cc.u32", "madc.lo.cc.u32", "madc.lo.u32"):
                t = ((v[0] * v[1]) & MASK) + v[2] + (cc if op.startswith("madc") else 0)
                r[dst] = t & MASK
                if ".cc" in op:
                    cc = t >> 32
            elif op in ("mad.hi.cc.u32", "madc.hi.cc.u32", "madc.hi.u32", "mad.hi.u32"):
                t = ((v[0] * v[1]) >> 32) + v[2] + (cc if op.startswith("madc") else 0)
                r[dst] = t & MASK
                if ".cc" in op:
                    cc = t >> 32
                else:
                    assert t >> 32 == 0, "dropped carry"
            elif op in ("add.cc.u32", "addc.cc.u32", "addc.u32", "add.u32"):
                t = v[0] + v[1] + (cc if op.startswith("addc") else 0)
                r[dst] = t & MASK
                if ".cc" in op:
                    cc = t >> 32
                else:
                    assert t >> 32 == 0, "dropped carry"
            elif op in ("sub.cc.u32", "subc.cc.u32", "subc.u32"):
                t = v[0] - v[1] - (cc if op.startswith("subc") else 0)
                r[dst] = t & MASK
                if ".cc" in op:
                    cc = 1 if t < 0 else 0
            elif op == "sub.u32":          # wrapping, leaves the carry flag alone
                r[dst] = (v[0] - v[1]) & MASK
            elif op == "min.u32":
                r[dst] = min(v[0], v[1])
            elif op == "and.b32":
                r[dst] = v[0] & v[1]
            elif op == "mov.u32":
                r[dst] = v[0]
            elif op == "setp.ne.u32":
                pred[dst] = v[0] != v[1]
            elif op == "selp.u32":
                r[dst] = v[0] if pred[srcs[2]] else v[1]
            else:
                raise ValueError(op)
        return r

    def count(self):
        wide = sum(1 for op, _, _ in self.ins if ".hi" in op)      # one IMAD.WIDE per lo/hi pair
        single = sum(1 for op, _, _ in self.ins if op == "mul.lo.u32" and True)
        return wide, len(self.ins)


def words(x, n):
    return [(x >> (32 * i)) & MASK for i in range(n)]


def emit_solinas_operands(g, m):
    """For modulus words equal to 0xffffffff:  m * 0xffffffff = (m - [m != 0]) * 2^32 + (-m mod 2^32), so the pair of
    a reduction chain takes two plain additions (ALU pipe) instead of one IMAD.WIDE (the busier fmaheavy pipe).
    The three instructions leave the carry flag alone (a fold carry may be pending)."""
    g.emit("sub.u32", "negm", 0, m)
    g.emit("sub.u32", "mt", m, 1)
    g.emit("min.u32", "mm1", "mt", m)


def solinas_words(mod, n, limit):
    """Indices of the 0xffffffff words of `mod` handled in add form: at most `limit` of them, lowest first."""
    idx = [j for j, w in enumerate(words(mod, n)) if w == MASK]
    return set(idx[:limit])


def gen_mul(n, mod, repl=frozenset()):
    """Instruction stream computing r = a*b*2^(-32n) mod `mod` (a, b < mod). Registers a0.., b0.. in, r0.. out.
    repl: word indices of the modulus (value 0xffffffff) reduced in add form (emit_solinas_operands)."""
    P = words(mod, n)
    m0 = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    g = Prog()
    live = set()
    E = lambda k: f"e{k}"
    O = lambda k: f"o{k}"

    def addend(reg):
        return reg if reg in live else 0

    for i in range(n):
        X, Y = (E, O) if i % 2 == 0 else (O, E)
        bi = f"b{i}"
        if i == 0:
            # row 0: every pair is fresh, the products are independent (no carries, no carry words)
            for j in range(n):
                F_ = X if j % 2 == 0 else Y
                g.emit("mul.lo.u32", F_(j), f"a{j}", bi)
                g.emit("mul.hi.u32", F_(j + 1), f"a{j}", bi)
                live.update((F_(j), F_(j + 1)))
        else:
            # ---- product chain A: even j, file X, pairs (i+j, i+j+1); carry-out into the carry word X[i+n]
            for j in range(0, n, 2):
                lo, hi = X(i + j), X(i + j + 1)
                assert lo in live and hi in live
                g.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", lo, f"a{j}", bi, lo)
                g.emit("madc.hi.cc.u32", hi, f"a{j}", bi, hi)
            top = X(i + n)
            assert top in live
            g.emit("addc.u32", top, top, 0)
            # ---- fold the other file's word at position i, its carry feeds product chain B (odd j, file Y)
            assert Y(i) in live and X(i) in live
            g.emit("add.cc.u32", X(i), X(i), Y(i))
            live.discard(Y(i))
            for j in range(1, n, 2):
                lo, hi = Y(i + j), Y(i + j + 1)
                assert lo in live
                g.emit("madc.lo.cc.u32", lo, f"a{j}", bi, lo)
                if j == n - 1:
                    assert hi not in live
                    g.emit("madc.hi.u32", hi, f"a{j}", bi, 0)   # fresh high word: cannot overflow
                else:
                    assert hi in live
                    g.emit("madc.hi.cc.u32", hi, f"a{j}", bi, hi)
                live.add(hi)
        # ---- Montgomery quotient
        if m0 == 1:
            m = X(i)
        else:
            m = "m"
            g.emit("mul.lo.u32", m, X(i), m0)
        if repl:
            emit_solinas_operands(g, m)
        # ---- reduction chain A: even j, file X.  The low word at position i becomes 0 and dies.
        started = False
        for j in range(0, n, 2):
            lo, hi = X(i + j), X(i + j + 1)
            pj = P[j]
            dst_lo = "junk" if j == 0 else lo        # position i: provably 0 afterwards (and m may alias X[i])
            if pj == 0:
                if started:
                    g.emit("addc.cc.u32", lo, addend(lo), 0)
                    g.emit("addc.cc.u32", hi, addend(hi), 0)
                    live.update((lo, hi))
                continue
            if pj == 1:
                g.emit("addc.cc.u32" if started else "add.cc.u32", dst_lo, addend(lo), m)
                g.emit("addc.cc.u32", hi, addend(hi), 0)
            elif pj == MASK and j in repl:
                g.emit("addc.cc.u32" if started else "add.cc.u32", dst_lo, addend(lo), "negm")
                g.emit("addc.cc.u32", hi, addend(hi), "mm1")
            else:
                g.emit("madc.lo.cc.u32" if started else "mad.lo.cc.u32", dst_lo, m, pj, addend(lo))
                g.emit("madc.hi.cc.u32", hi, m, pj, addend(hi))
            live.update((lo, hi))
            started = True
        assert started, "p[0] is odd, so chain A always starts"
        live.discard(X(i))
        top = X(i + n)
        g.emit("addc.u32", top, addend(top), 0)
        live.add(top)
        # ---- reduction chain B: odd j, file Y; carry-out into the fresh carry word Y[i+n+1]
        started = False
        for j in range(1, n, 2):
            lo, hi = Y(i + j), Y(i + j + 1)
            pj = P[j]
            if pj == 0:
                if started:
                    g.emit("addc.cc.u32", lo, addend(lo), 0)
                    g.emit("addc.cc.u32", hi, addend(hi), 0)
                    live.update((lo, hi))
                continue
            if pj == 1:
                g.emit("addc.cc.u32" if started else "add.cc.u32", lo, addend(lo), m)
                g.emit("addc.cc.u32", hi, addend(hi), 0)
            elif pj == MASK and j in repl:
                g.emit("addc.cc.u32" if started else "add.cc.u32", lo, addend(lo), "negm")
                g.emit("addc.cc.u32", hi, addend(hi), "mm1")
            else:
                g.emit("madc.lo.cc.u32" if started else "mad.lo.cc.u32", lo, m, pj, addend(lo))
                g.emit("madc.hi.cc.u32", hi, m, pj, addend(hi))
            live.update((lo, hi))
            started = True
        top = Y(i + n + 1)
        assert top not in live
        if started:
            g.emit("addc.u32", top, 0, 0)
        else:
            g.emit("mov.u32", top, 0)
        live.add(top)
    # ---- merge positions n .. 2n (+1) of both files into t0..tn
    first = True
    for k in range(n):
        ek, ok = E(n + k), O(n + k)
        srcs = [r for r in (ek, ok) if r in live]
        a0 = srcs[0] if srcs else 0
        a1 = srcs[1] if len(srcs) > 1 else 0
        g.emit("add.cc.u32" if first else "addc.cc.u32", f"t{k}", a0, a1)
        first = False
    tops = [r for r in (E(2 * n), O(2 * n)) if r in live]
    assert E(2 * n + 1) not in live and O(2 * n + 1) not in live
    g.emit("addc.u32", f"t{n}", tops[0] if tops else 0, tops[1] if len(tops) > 1 else 0)
    # ---- conditional subtraction: r = t - p if t >= p else t   (t < 2p)
    for k in range(n):
        g.emit("sub.cc.u32" if k == 0 else "subc.cc.u32", f"d{k}", f"t{k}", P[k])
    g.emit("subc.u32", "dt", f"t{n}", 0)           # 0xffffffff iff t < p
    g.emit("setp.ne.u32", "pq", "dt", 0)
    for k in range(n):
        g.emit("selp.u32", f"r{k}", f"t{k}", f"d{k}", "pq")
    return g


def gen_sqr(n, mod, repl=frozenset()):
    """r = a*a*2^(-32n) mod `mod`.  The product phase uses the symmetry of squaring: the n(n-1)/2 cross products
    a_i*a_j (i < j) are accumulated once, doubled by a one-bit shift, and the n squares a_i^2 are added — n(n+1)/2
    wide multiplies instead of n^2.  The double-width product then sits in file E (even-aligned pairs); the n
    reduction rows run as in gen_mul with file O collecting the odd-position products.  Because E's upper words are
    full product words, carry-outs of chains that end in E are collected in separate registers k<pos> (each at most
    2) and added after the final merge."""
    P = words(mod, n)
    m0 = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    g = Prog()
    live = set()
    E = lambda k: f"e{k}"
    O = lambda k: f"o{k}"
    K = lambda k: f"k{k}"

    def addend(reg):
        return reg if reg in live else 0

    # ---- phase 1: cross products sum_{i<j} a_i a_j 2^(32(i+j)); position i+j even -> file E, odd -> file O
    for i in range(n - 1):
        for parity in (0, 1):
            F = E if parity == 0 else O
            js = [j for j in range(i + 1, n) if (i + j) % 2 == parity]
            if not js:
                continue
            fresh_all = all(F(i + j) not in live and F(i + j + 1) not in live for j in js)
            if fresh_all:
                for j in js:
                    g.emit("mul.lo.u32", F(i + j), f"a{i}", f"a{j}")
                    g.emit("mul.hi.u32", F(i + j + 1), f"a{i}", f"a{j}")
                    live.update((F(i + j), F(i + j + 1)))
                continue
            for idx, j in enumerate(js):
                lo, hi = F(i + j), F(i + j + 1)
                last = idx == len(js) - 1
                g.emit("mad.lo.cc.u32" if idx == 0 else "madc.lo.cc.u32", lo, f"a{i}", f"a{j}", addend(lo))
                if last and hi not in live:
                    g.emit("madc.hi.u32", hi, f"a{i}", f"a{j}", 0)          # fresh high word: cannot overflow
                    live.update((lo, hi))
                else:
                    g.emit("madc.hi.cc.u32", hi, f"a{i}", f"a{j}", addend(hi))
                    live.update((lo, hi))
                    if last:
                        top = F(i + j + 2)
                        assert top not in live
                        g.emit("addc.u32", top, 0, 0)                        # carry word for the next row's top pair
                        live.add(top)
    # ---- merge O into E (positions 1 .. 2n-1), then double (shift left by one bit)
    first = True
    for k in range(1, 2 * n):
        srcs = [r for r in (E(k), O(k)) if r in live]
        if not srcs:
            g.emit("addc.u32" if not first else "mov.u32", E(k), *([0, 0] if not first else [0]))
            live.add(E(k))
            continue
        a0 = srcs[0]
        a1 = srcs[1] if len(srcs) > 1 else 0
        if k == 2 * n - 1:
            g.emit("add.u32" if first else "addc.u32", E(k), a0, a1)
        else:
            g.emit("add.cc.u32" if first else "addc.cc.u32", E(k), a0, a1)
        first = False
        live.add(E(k))
    for r in list(live):
        if r.startswith("o"):
            live.discard(r)
    for k in range(1, 2 * n):
        if k == 2 * n - 1:
            g.emit("add.u32" if k == 1 else "addc.u32", E(k), E(k), E(k))
        else:
            g.emit("add.cc.u32" if k == 1 else "addc.cc.u32", E(k), E(k), E(k))
    # ---- squares a_i^2 at position 2i: one chain over the even-aligned pairs of E
    g.emit("mul.lo.u32", E(0), "a0", "a0")
    g.emit("mad.hi.cc.u32", E(1), "a0", "a0", E(1))
    live.add(E(0))
    for i in range(1, n):
        g.emit("madc.lo.cc.u32", E(2 * i), f"a{i}", f"a{i}", E(2 * i))
        if i == n - 1:
            g.emit("madc.hi.u32", E(2 * i + 1), f"a{i}", f"a{i}", E(2 * i + 1))   # a^2 < 2^(64n): no carry out
        else:
            g.emit("madc.hi.cc.u32", E(2 * i + 1), f"a{i}", f"a{i}", E(2 * i + 1))
    # ---- phase 3: n reduction rows
    klive = set()
    for i in range(n):
        X, Y = (E, O) if i % 2 == 0 else (O, E)
        x_is_e = (i % 2 == 0)
        if i > 0:
            # fold; carry feeds chain B (it starts at position i+1 in Y).  A word that was never written is 0
            # (moduli with zero words, e.g. P-384, leave gaps in file O).
            g.emit("add.cc.u32", X(i), addend(X(i)), addend(Y(i)))
            live.add(X(i))
            live.discard(Y(i))
        if m0 == 1:
            m = X(i)
        else:
            m = "m"
            g.emit("mul.lo.u32", m, X(i), m0)
        if repl:
            emit_solinas_operands(g, m)
        # chain B first when there is a fold carry pending (i > 0); chain A afterwards (it starts its own chain)
        def chain(F, js, carry_in, f_is_e, top_pos):
            started = carry_in
            for j in js:
                lo, hi = F(i + j), F(i + j + 1)
                pj = P[j]
                is_pos_i = (j == 0)
                dst_lo = "junk" if is_pos_i else lo
                if pj == 0:
                    if started:
                        g.emit("addc.cc.u32", lo, addend(lo), 0)
                        g.emit("addc.cc.u32", hi, addend(hi), 0)
                        live.update((lo, hi))
                    continue
                if pj == 1:
                    g.emit("addc.cc.u32" if started else "add.cc.u32", dst_lo, addend(lo), m)
                    g.emit("addc.cc.u32", hi, addend(hi), 0)
                elif pj == MASK and j in repl:
                    g.emit("addc.cc.u32" if started else "add.cc.u32", dst_lo, addend(lo), "negm")
                    g.emit("addc.cc.u32", hi, addend(hi), "mm1")
                else:
                    g.emit("madc.lo.cc.u32" if started else "mad.lo.cc.u32", dst_lo, m, pj, addend(lo))
                    g.emit("madc.hi.cc.u32", hi, m, pj, addend(hi))
                live.update((lo, hi))
                started = True
            # carry-out at position top_pos
            if f_is_e:
                kreg = K(top_pos)
                if started:
                    g.emit("addc.u32", kreg, kreg if kreg in klive else 0, 0)
                    klive.add(kreg)
            else:
                top = F(top_pos)
                if started:
                    g.emit("addc.u32", top, addend(top), 0)
                    live.add(top)
        if i > 0:
            chain(Y, list(range(1, n, 2)), True, not x_is_e, i + n + 1)
            chain(X, list(range(0, n, 2)), False, x_is_e, i + n)
        else:
            chain(X, list(range(0, n, 2)), False, x_is_e, i + n)
            chain(Y, list(range(1, n, 2)), False, not x_is_e, i + n + 1)
        live.discard(X(i))
    # ---- final: t = E[n..2n-1] + O[n..2n-1] (+ carry into t_n), then t += K
    first = True
    for k in range(n):
        srcs = [r for r in (E(n + k), O(n + k)) if r in live]
        a0 = srcs[0] if srcs else 0
        a1 = srcs[1] if len(srcs) > 1 else 0
        g.emit("add.cc.u32" if first else "addc.cc.u32", f"t{k}", a0, a1)
        first = False
    tops = [r for r in (E(2 * n), O(2 * n)) if r in live]
    g.emit("addc.u32", f"t{n}", tops[0] if tops else 0, tops[1] if len(tops) > 1 else 0)
    first = True
    for k in range(n + 1):
        kreg = K(n + k)
        src = kreg if kreg in klive else 0
        if k == n:
            g.emit("add.u32" if first else "addc.u32", f"t{k}", f"t{k}", src)
        else:
            g.emit("add.cc.u32" if first else "addc.cc.u32", f"t{k}", f"t{k}", src)
        first = False
    for k in range(n):
        g.emit("sub.cc.u32" if k == 0 else "subc.cc.u32", f"d{k}", f"t{k}", P[k])
    g.emit("subc.u32", "dt", f"t{n}", 0)
    g.emit("setp.ne.u32", "pq", "dt", 0)
    for k in range(n):
        g.emit("selp.u32", f"r{k}", f"t{k}", f"d{k}", "pq")
    return g


def gen_add(n, mod):
    P = words(mod, n)
    g = Prog()
    for k in range(n):
        g.emit("add.cc.u32" if k == 0 else "addc.cc.u32", f"t{k}", f"a{k}", f"b{k}")
    g.emit("addc.u32", "tc", 0, 0)
    for k in range(n):
        g.emit("sub.cc.u32" if k == 0 else "subc.cc.u32", f"d{k}", f"t{k}", P[k])
    g.emit("subc.u32", "dt", "tc", 0)
    g.emit("setp.ne.u32", "pq", "dt", 0)
    for k in range(n):
        g.emit("selp.u32", f"r{k}", f"t{k}", f"d{k}", "pq")
    return g


def gen_sub(n, mod):
    P = words(mod, n)
    g = Prog()
    for k in range(n):
        g.emit("sub.cc.u32" if k == 0 else "subc.cc.u32", f"t{k}", f"a{k}", f"b{k}")
    g.emit("subc.u32", "bm", 0, 0)                 # 0xffffffff iff a < b
    for k in range(n):
        if P[k] == 0:
            g.emit("mov.u32", f"q{k}", 0)
        else:
            g.emit("and.b32", f"q{k}", "bm", P[k])
    for k in range(n):
        g.emit("add.cc.u32" if k == 0 else "addc.cc.u32", f"r{k}", f"t{k}", f"q{k}")
    return g


# ---- rendering ---------------------------------------------------------------------------------------------

def render_asm(prog: Prog, n: int, two_inputs=True, indent="\t\t"):
    """C++ asm statement: outputs r0..r{n-1} = %0..%{n-1}; inputs a = %n.., b = %2n.."""
    opmap = {}
    for k in range(n):
        opmap[f"r{k}"] = f"%{k}"
        opmap[f"a{k}"] = f"%{n + k}"
        opmap[f"b{k}"] = f"%{2 * n + k}"
    temps = sorted(r for r in prog.regs if r not in opmap and r != "pq")
    lines = ["{"]
    lines.append(".reg .u32 " + ", ".join(temps) + ";")
    lines.append(".reg .pred pq;")

    def fmt(s):
        if isinstance(s, int):
            return "0x%08x" % s if s > 9 else str(s)
        return opmap.get(s, s)

    for op, dst, srcs in prog.ins:
        lines.append(f"{op} {fmt(dst)}, " + ", ".join(fmt(s) for s in srcs) + ";")
    lines.append("}")
    body = ("\n" + indent + "    ").join('"' + ln + '\\n\\t"' for ln in lines)
    outs = ", ".join(f'"=r"(r.w[{k}])' for k in range(n))
    ins = ", ".join(f'"r"(a.w[{k}])' for k in range(n))
    if two_inputs:
        ins += ", " + ", ".join(f'"r"(b.w[{k}])' for k in range(n))
    return f"{indent}asm({body}\n{indent}    : {outs}\n{indent}    : {ins});"


def fields():
    for name, (cid, p, a, b, q, gx, gy) in CURVES.items():
        n = nwords(p)
        yield f"Fp_{name}", n, p
        yield f"Fq_{name}", n, q


# Share of a modulus' 0xffffffff words reduced in add form, per multiplier variant (the inlined multiplier of the
# fixed-base kernels / the out-of-line one of K2 and K3).  MEASURED SLOWER on B200 and therefore OFF by default:
# with shares (0.5, 1.0) secp256r1 fixed base fell from 478 to 449 M/s, ECDSA verify from 20.2 to 18.5 M/s, secp384r1
# from 149 to 144 M/s, secp521r1 from 37.6 to 35.5 M/s; with (1.0, 1.0) further (436 / 18.5 / 136 / 32.1).  One
# IMAD.WIDE (fmaheavy pipe, issue rate 1/4) is replaced by two IADD3.X on the ALU pipe (issue rate 1/2 each), which
# is cycle-neutral at best, plus three ALU instructions per row, and the dependent carry chains get longer.
# Kept as a generator option (and covered by tests/test_ptx_emulation.py): ECC_SOLINAS_INLINE / ECC_SOLINAS_CALL.
SOLINAS_SHARE = {"inline": float(os.environ.get("ECC_SOLINAS_INLINE", "0")),
                 "call": float(os.environ.get("ECC_SOLINAS_CALL", "0"))}


def variant_repl(mod, n, variant):
    cnt = sum(1 for w in words(mod, n) if w == MASK)
    return frozenset(solinas_words(mod, n, int(cnt * SOLINAS_SHARE[variant] + 1e-9)))


def variants(n, mod):
    """[(variant name, mul program, sqr program)]; one entry when both variants are the same program."""
    ri, rc = variant_repl(mod, n, "inline"), variant_repl(mod, n, "call")
    if ri == rc:
        return [("both", gen_mul(n, mod, ri), gen_sqr(n, mod, ri))]
    return [("inline", gen_mul(n, mod, ri), gen_sqr(n, mod, ri)), ("call", gen_mul(n, mod, rc), gen_sqr(n, mod, rc))]


def main():
    out = ["/* GENERATED by tools/gen_fp_ptx.py — do not edit.  Inline-PTX field arithmetic (device only). */",
           "#pragma once", "",
           "#if defined(ECC_INLINE_MUL)", "#define ECC_MUL_LINKAGE __forceinline__", "#else",
           "#define ECC_MUL_LINKAGE __noinline__", "#endif",
           "/* the squaring can be kept out of line on its own (instruction-cache experiments on K1, DESIGN.md §4) */",
           "#if defined(ECC_INLINE_MUL) && !defined(ECC_NOINLINE_SQR)", "#define ECC_SQR_LINKAGE __forceinline__", "#else",
           "#define ECC_SQR_LINKAGE __noinline__", "#endif", "", "namespace eccb200 {", "",
           "template <class F> struct FieldPtx;", ""]
    for tag, n, mod in fields():
        add, sub = gen_add(n, mod), gen_sub(n, mod)
        vs = variants(n, mod)
        out.append(f"template <> struct FieldPtx<{tag}> {{")
        out.append(f"\tstatic constexpr int N = {n};")
        out.append(f"\ttypedef Fe<{n}> E;")
        out.append("\t/* Out-of-line, operands and result BY VALUE (they stay in registers; ptxas allocates across the call):")
        out.append("\t * one copy of the ~200-instruction product per kernel instead of one per use keeps the hot loops of")
        out.append("\t * the scalar-multiplication kernels inside the instruction cache (profiles/: no_instruction stalls). */")
        for vi, (vname, mul, sqr) in enumerate(vs):
            if len(vs) > 1:
                out.append("#if defined(ECC_INLINE_MUL)" if vi == 0 else "#else")
            wide, total = mul.count()
            swide, stotal = sqr.count()
            out.append(f"\t/* {tag} [{vname}]: mul = {wide} wide multiply-accumulates, {total} PTX instructions; "
                       f"sqr = {swide} wide, {stotal} PTX instructions */")
            out.append("\tstatic __device__ ECC_MUL_LINKAGE E mul_fn(E a, E b)\n\t{\n\t\tE r;")
            out.append(render_asm(mul, n))
            out.append("\t\treturn r;\n\t}")
            out.append("\tstatic __device__ ECC_SQR_LINKAGE E sqr_fn(E a)\n\t{\n\t\tE r;")
            out.append(render_asm(sqr, n, two_inputs=False))
            out.append("\t\treturn r;\n\t}")
        if len(vs) > 1:
            out.append("#endif")
        out.append("\tstatic __device__ __forceinline__ void mul(E &r, const E &a, const E &b) { r = mul_fn(a, b); }")
        out.append("\tstatic __device__ __forceinline__ void sqr(E &r, const E &a) { r = sqr_fn(a); }")
        out.append("\tstatic __device__ __forceinline__ void add(E &r, const E &a, const E &b)\n\t{")
        out.append(render_asm(add, n))
        out.append("\t}")
        out.append("\tstatic __device__ __forceinline__ void sub(E &r, const E &a, const E &b)\n\t{")
        out.append(render_asm(sub, n))
        out.append("\t}")
        out.append("};")
        out.append("")
    out.append("} // namespace eccb200")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "libecc_b200", "csrc", "fp_ptx.cuh")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
