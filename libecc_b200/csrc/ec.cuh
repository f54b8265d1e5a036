/*
 * ec.cuh — short-Weierstrass group law on the device (replaces the reference's src/curves/prj_pt.c hot path).
 *
 * The reference computes prj_pt_mul with a masked Montgomery ladder over the Renes-Costello-Batina complete
 * addition in homogeneous projective coordinates (curves/prj_pt.c:971-1071, :1569-1720): 513 complete additions
 * = 8 724 field multiplications per 256-bit scalar.  Its result is only defined up to the projective class (the
 * input is blinded with a random lambda, :1266-1291), i.e. by the affine point.  The device is therefore free
 * to use a different algorithm as long as the affine output / infinity flag / error code agree:
 *
 *   - Jacobian coordinates (X/Z^2, Y/Z^3), a = -3 doubling (dbl-2001-b, 3M+5S), general add (12M+4S) and mixed
 *     add with an affine operand (8M+3S);
 *   - every exceptional case of the incomplete formulas is handled explicitly (P = inf, Q = inf, P = Q -> double,
 *     P = -Q -> inf), so the result is the group-law result for ALL inputs, like the reference's complete formulas;
 *   - fixed base (k*G): comb over a precomputed affine table T[i][d] = d * 2^(w*i) * G, one mixed add per
 *     window, no doublings (K1);
 *   - variable base (k*P): signed 4-bit fixed window, 8-entry Jacobian table per thread (K2).
 *
 * The three curves of BASELINE.json have a = p - 3 (curves/known/ec_params_secp256r1.h:78-83, ..._frp256v1.h:84-89,
 * ..._secp384r1.h) and use the a = -3 doubling; the additional curves (Brainpool P256r1 / P384r1: generic a,
 * secp256k1: a = 0) select their doubling through Curve::A_KIND (tools/gen_curve_constants.py).
 */
#pragma once
#include "fp.cuh"

namespace eccb200 {

template <class C> struct Jac {
	Fe<C::N> X, Y, Z; /* Z == 0 <=> point at infinity (prj_pt_iszero, curves/prj_pt.c:107) */
};

template <class C> struct Aff {
	Fe<C::N> x, y; /* Montgomery form */
};

template <class C> struct EC {
	static constexpr int N = C::N;
	typedef Field<typename C::Fp> F;
	typedef Fe<N> E;
	typedef Jac<C> J;
	typedef Aff<C> A;

	static ECC_HD void set_inf(J &p)
	{
		F::set_one(p.X);
		F::set_one(p.Y);
		F::set_zero(p.Z);
	}
	static ECC_HD bool is_inf(const J &p) { return F::is_zero(p.Z); }

	static ECC_HD void from_affine(J &p, const A &a)
	{
		p.X = a.x;
		p.Y = a.y;
		F::set_one(p.Z);
	}

	static ECC_HD void load_a(E &r)
	{
#pragma unroll
		for (int i = 0; i < N; i++) r.w[i] = C::A_MONT(i);
	}

	/* u = x^3 + a x + b (Montgomery form): the right-hand side of the curve equation */
	static ECC_HD void curve_rhs(E &u, const E &x)
	{
		E t, b;
		F::sqr(t, x);
		F::mul(u, t, x);      /* x^3 */
		if (C::A_KIND == 0) {
			F::add(t, x, x);
			F::add(t, t, x);  /* 3x */
			F::sub(u, u, t);
		} else if (C::A_KIND == 2) {
			E am;
			load_a(am);
			F::mul(t, am, x);
			F::add(u, u, t);
		}
#pragma unroll
		for (int i = 0; i < N; i++) b.w[i] = C::B_MONT(i);
		F::add(u, u, b);
	}

	/* y^2 == x^3 + a x + b (all Montgomery form); affine form of prj_pt_is_on_curve (curves/prj_pt.c:144-190) */
	static ECC_HD bool on_curve(const A &a)
	{
		E t, u;
		curve_rhs(u, a.x);
		F::sqr(t, a.y);
		return F::eq(t, u);
	}

	/* Out-of-line copy of dbl for the exceptional (P == Q) branches of the additions: keeps the rarely taken
	 * path from being inlined into every hot loop. */
	static ECC_NOINLINE void dbl_slow(J &r, const J &p) { dbl(r, p); }

	/* Jacobian doubling, r may alias p, inf -> inf (Z3 = 0).  Three formulas selected at compile time by the curve:
	 * a = -3 (dbl-2001-b, 3M + 5S: the three curves of BASELINE.json), a = 0 (dbl-2009-l, 2M + 5S: secp256k1) and
	 * generic a (dbl-2007-bl, 2M + 8S: Brainpool). */
	static ECC_HD void dbl(J &r, const J &p)
	{
		if (C::A_KIND == 1) {
			E a_, b_, c_, d_, e_, f_, t;
			F::sqr(a_, p.X);
			F::sqr(b_, p.Y);
			F::sqr(c_, b_);
			F::add(t, p.X, b_);
			F::sqr(d_, t);
			F::sub(d_, d_, a_);
			F::sub(d_, d_, c_);
			F::add(d_, d_, d_);      /* D = 2((X+B)^2 - A - C) */
			F::add(e_, a_, a_);
			F::add(e_, e_, a_);      /* E = 3A */
			F::sqr(f_, e_);
			F::mul(t, p.Y, p.Z);
			F::add(r.Z, t, t);       /* Z3 = 2YZ */
			F::sub(f_, f_, d_);
			F::sub(r.X, f_, d_);     /* X3 = F - 2D */
			F::sub(t, d_, r.X);
			F::mul(f_, e_, t);
			F::add(c_, c_, c_);
			F::add(c_, c_, c_);
			F::add(c_, c_, c_);      /* 8C */
			F::sub(r.Y, f_, c_);
			return;
		}
		if (C::A_KIND == 2) {
			E xx, yy, yyyy, zz, s_, m_, t, am;
			F::sqr(xx, p.X);
			F::sqr(yy, p.Y);
			F::sqr(yyyy, yy);
			F::sqr(zz, p.Z);
			F::add(t, p.X, yy);
			F::sqr(s_, t);
			F::sub(s_, s_, xx);
			F::sub(s_, s_, yyyy);
			F::add(s_, s_, s_);      /* S = 2((X+YY)^2 - XX - YYYY) */
			F::add(t, p.Y, p.Z);
			F::sqr(r.Z, t);
			F::sub(r.Z, r.Z, yy);
			F::sub(r.Z, r.Z, zz);    /* Z3 = (Y+Z)^2 - YY - ZZ */
			F::sqr(t, zz);
			load_a(am);
			F::mul(m_, am, t);       /* a ZZ^2 */
			F::add(t, xx, xx);
			F::add(t, t, xx);
			F::add(m_, m_, t);       /* M = 3XX + a ZZ^2 */
			F::sqr(t, m_);
			F::sub(t, t, s_);
			F::sub(r.X, t, s_);      /* X3 = M^2 - 2S */
			F::sub(t, s_, r.X);
			F::mul(s_, m_, t);
			F::add(yyyy, yyyy, yyyy);
			F::add(yyyy, yyyy, yyyy);
			F::add(yyyy, yyyy, yyyy); /* 8 YYYY */
			F::sub(r.Y, s_, yyyy);
			return;
		}
		E delta, gamma, beta, alpha, t0, t1;
		F::sqr(delta, p.Z);
		F::sqr(gamma, p.Y);
		F::mul(beta, p.X, gamma);
		F::sub(t0, p.X, delta);
		F::add(t1, p.X, delta);
		F::mul(alpha, t0, t1);
		F::add(t0, alpha, alpha);
		F::add(alpha, t0, alpha); /* 3 (X-delta)(X+delta) */
		F::add(t1, p.Y, p.Z);
		F::sqr(t0, t1);
		F::sub(t0, t0, gamma);
		F::sub(r.Z, t0, delta);   /* Z3 = (Y+Z)^2 - gamma - delta */
		F::add(t0, beta, beta);
		F::add(t0, t0, t0);       /* 4 beta */
		F::add(t1, t0, t0);       /* 8 beta */
		F::sqr(r.X, alpha);
		F::sub(r.X, r.X, t1);     /* X3 = alpha^2 - 8 beta */
		F::sub(t0, t0, r.X);
		F::mul(t1, alpha, t0);
		F::sqr(t0, gamma);
		F::add(t0, t0, t0);
		F::add(t0, t0, t0);
		F::add(t0, t0, t0);       /* 8 gamma^2 */
		F::sub(r.Y, t1, t0);
	}

	/*
	 * r = p + q, q affine (Z2 = 1): 8M + 3S.  Exceptional cases resolved explicitly so that the result equals the
	 * group law for every input, matching what prj_pt_add's complete formulas give (curves/prj_pt.c:1204).
	 */
	static ECC_HD void add_mixed(J &r, const J &p, const A &q)
	{
		E z1z1, u2, s2, h, rr, hh, hhh, v, t;
		F::sqr(z1z1, p.Z);
		F::mul(u2, q.x, z1z1);
		F::mul(t, q.y, p.Z);
		F::mul(s2, t, z1z1);
		F::sub(h, u2, p.X);
		F::sub(rr, s2, p.Y);
		bool p_inf = F::is_zero(p.Z);
		bool h0 = F::is_zero(h);
		if (p_inf) {
			from_affine(r, q);
			return;
		}
		if (h0) {
			if (F::is_zero(rr)) {
				J qq;
				from_affine(qq, q);
				dbl_slow(r, qq);
			} else {
				set_inf(r);
			}
			return;
		}
		F::sqr(hh, h);
		F::mul(hhh, h, hh);
		F::mul(v, p.X, hh);
		E x3, y3, z3;
		F::sqr(x3, rr);
		F::sub(x3, x3, hhh);
		F::sub(x3, x3, v);
		F::sub(x3, x3, v);
		F::sub(t, v, x3);
		F::mul(y3, rr, t);
		F::mul(t, p.Y, hhh);
		F::sub(y3, y3, t);
		F::mul(z3, p.Z, h);
		r.X = x3;
		r.Y = y3;
		r.Z = z3;
	}

	/* r = p + q, both Jacobian (add-1998-cmo-2): 12M + 4S, exceptional cases resolved explicitly. */
	static ECC_HD void add_full(J &r, const J &p, const J &q)
	{
		E z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t;
		F::sqr(z1z1, p.Z);
		F::sqr(z2z2, q.Z);
		F::mul(u1, p.X, z2z2);
		F::mul(u2, q.X, z1z1);
		F::mul(t, p.Y, q.Z);
		F::mul(s1, t, z2z2);
		F::mul(t, q.Y, p.Z);
		F::mul(s2, t, z1z1);
		F::sub(h, u2, u1);
		F::sub(rr, s2, s1);
		bool p_inf = F::is_zero(p.Z), q_inf = F::is_zero(q.Z);
		if (p_inf) {
			r = q;
			return;
		}
		if (q_inf) {
			r = p;
			return;
		}
		if (F::is_zero(h)) {
			if (F::is_zero(rr)) {
				J pp = p;
				dbl_slow(r, pp);
			} else {
				set_inf(r);
			}
			return;
		}
		F::sqr(hh, h);
		F::mul(hhh, h, hh);
		F::mul(v, u1, hh);
		E x3, y3, z3;
		F::sqr(x3, rr);
		F::sub(x3, x3, hhh);
		F::sub(x3, x3, v);
		F::sub(x3, x3, v);
		F::sub(t, v, x3);
		F::mul(y3, rr, t);
		F::mul(t, s1, hhh);
		F::sub(y3, y3, t);
		F::mul(t, p.Z, q.Z);
		F::mul(z3, t, h);
		r.X = x3;
		r.Y = y3;
		r.Z = z3;
	}

	/* out-of-line copy of add_full for code that adds up many points outside the hot loops (bucket reduction of the
	 * multi-scalar multiplication): one body per kernel instead of one per call site */
	static ECC_NOINLINE void add_full_ool(J &r, const J &p, const J &q) { add_full(r, p, q); }

	static ECC_HD void neg(J &r, const J &p)
	{
		r.X = p.X;
		r.Z = p.Z;
		F::neg(r.Y, p.Y);
	}

	/*
	 * Extended Jacobian ("XYZZ") accumulator of the fixed-base comb: x = X/ZZ, y = Y/ZZZ with ZZ^3 == ZZZ^2;
	 * ZZ == 0 <=> point at infinity.  Mixed addition madd-2008-s: 8M + 2S (one squaring less than the Jacobian
	 * mixed addition, because Z^2 and Z^3 are carried instead of recomputed).  Exceptional cases as in add_mixed.
	 */
	struct XZ {
		E X, Y, ZZ, ZZZ;
	};
	static ECC_HD void xz_set_inf(XZ &p)
	{
		F::set_one(p.X);
		F::set_one(p.Y);
		F::set_zero(p.ZZ);
		F::set_zero(p.ZZZ);
	}
	static ECC_HD void xz_from_affine(XZ &p, const A &a)
	{
		p.X = a.x;
		p.Y = a.y;
		F::set_one(p.ZZ);
		F::set_one(p.ZZZ);
	}
	/* rarely taken P == Q branch of xz_add_mixed: 2Q through the Jacobian doubling, then ZZ = Z^2, ZZZ = Z^3 */
	static ECC_NOINLINE void xz_dbl_affine_slow(XZ &r, const A &q)
	{
		J qq, d;
		from_affine(qq, q);
		dbl(d, qq);
		r.X = d.X;
		r.Y = d.Y;
		F::sqr(r.ZZ, d.Z);
		F::mul(r.ZZZ, r.ZZ, d.Z);
	}
	/* the i-th of the eight products of xz_add_mixed: the first ECC_K1_OOL_MULS of them call the out-of-line copy when
	 * the translation unit inlines the multiplier (K1), see Field::mul_ool */
#ifndef ECC_K1_OOL_MULS
#define ECC_K1_OOL_MULS 0
#endif
	template <int I> static ECC_HD void xz_mul(E &r, const E &a, const E &b)
	{
#if defined(ECC_INLINE_MUL)
		if (I < ECC_K1_OOL_MULS) {
			F::mul_ool(r, a, b);
			return;
		}
#endif
		F::mul(r, a, b);
	}
	static ECC_HD void xz_add_mixed(XZ &r, const XZ &p, const A &q)
	{
		E u2, s2, pp_, rr, ppp, qv, t;
		if (F::is_zero(p.ZZ)) { /* first non-zero window of the comb */
			xz_from_affine(r, q);
			return;
		}
		xz_mul<0>(u2, q.x, p.ZZ);
		xz_mul<1>(s2, q.y, p.ZZZ);
		F::sub(pp_, u2, p.X); /* P */
		F::sub(rr, s2, p.Y);  /* R */
		if (F::is_zero(pp_)) {
			if (F::is_zero(rr)) xz_dbl_affine_slow(r, q);
			else xz_set_inf(r);
			return;
		}
		E x3, y3;
		F::sqr(t, pp_);        /* PP */
		xz_mul<2>(ppp, pp_, t);   /* PPP */
		xz_mul<3>(qv, p.X, t);    /* Q = X1 * PP */
		E zz3;
		xz_mul<4>(zz3, p.ZZ, t);  /* ZZ3 = ZZ1 * PP */
		F::sqr(x3, rr);
		F::sub(x3, x3, ppp);
		F::sub(x3, x3, qv);
		F::sub(x3, x3, qv);    /* X3 = R^2 - PPP - 2Q */
		F::sub(t, qv, x3);
		xz_mul<5>(y3, rr, t);
		xz_mul<6>(t, p.Y, ppp);
		F::sub(y3, y3, t);     /* Y3 = R (Q - X3) - Y1 PPP */
		xz_mul<7>(t, p.ZZZ, ppp); /* ZZZ3 = ZZZ1 * PPP */
		r.ZZZ = t;
		r.ZZ = zz3;
		r.X = x3;
		r.Y = y3;
	}
	/* XYZZ -> Jacobian with Z' = ZZ: x = X ZZ / ZZ^2, y = Y ZZZ / ZZ^3 (ZZ^3 == ZZZ^2).  2M. */
	static ECC_HD void xz_to_jac(J &r, const XZ &p)
	{
		F::mul(r.X, p.X, p.ZZ);
		F::mul(r.Y, p.Y, p.ZZZ);
		r.Z = p.ZZ;
	}
};

/* ---------------------------------------------------------------------------------------------- wire format */

/* len big-endian bytes (libecc wire format, nn_init_from_buf nn/nn.c:479) -> N little-endian 32-bit words; portable
 * byte-wise form for the host build of the tests (the kernels use load_wire / store_wire of kernels.cuh) */
template <int N> ECC_HD void load_be(Fe<N> &r, const uint8_t *buf, int len = 4 * N)
{
	for (int i = 0; i < N; i++) r.w[i] = 0;
	for (int j = 0; j < len && j < 4 * N; j++) r.w[j >> 2] |= (uint32_t)buf[len - 1 - j] << (8 * (j & 3));
}

/* nn_export_to_buf (nn/nn.c:511) */
template <int N> ECC_HD void store_be(uint8_t *buf, const Fe<N> &a, int len = 4 * N)
{
	for (int j = 0; j < len; j++) buf[len - 1 - j] = (j < 4 * N) ? (uint8_t)(a.w[j >> 2] >> (8 * (j & 3))) : 0;
}

/* ---------------------------------------------------------------------------------------------- scalars */

/* Reduce a raw wire scalar (QLEN bytes, so < 2^(8*QLEN)) modulo q: the reference's ladder yields (k mod q)*P for any
 * k (curves/prj_pt.c:1591-1619).  Shifted conditional subtractions of q << sh, sh = 8*QLEN - bitlen(q) .. 0: a single
 * step for the 256/384-bit curves (2^(8*QLEN) < 2q), eight for the 521-bit one (66-byte scalars, q < 2^521). */
template <class C> ECC_HD void scalar_reduce(Fe<C::N> &k)
{
	typedef Field<typename C::Fq> Fq;
	constexpr int N = C::N;
	constexpr int SH = 8 * C::QLEN - C::QBITS;
	static_assert(SH >= 0 && SH < 32 && C::QBITS + SH <= 32 * N, "scalar_reduce: q << SH must fit N words");
#pragma unroll
	for (int sh = SH; sh > 0; sh--) {
		uint32_t d[N];
		uint64_t bw = 0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint32_t lo = (i > 0) ? C::Fq::P(i - 1) : 0u;
			uint32_t qs = (C::Fq::P(i) << sh) | (lo >> (32 - sh)); /* word i of q << sh */
			uint64_t t = (uint64_t)k.w[i] - qs - bw;
			d[i] = (uint32_t)t;
			bw = (t >> 32) & 1;
		}
#pragma unroll
		for (int i = 0; i < N; i++) k.w[i] = bw ? k.w[i] : d[i];
	}
	for (int it = 0; it < 2; it++) Fq::cond_sub_mod(k, k);
}

/* 64-bit funnel shifts on word pairs (static register indices only: the scalar walks through the window loop by
 * being shifted, never by dynamic indexing — see DESIGN.md §8, "nvcc stack-colouring hazard"). */
ECC_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) /* low word of ((hi:lo) >> sh), 0 <= sh < 32 */
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(lo, hi, sh);
#else
	return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}
ECC_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, int sh) /* high word of ((hi:lo) << sh), 0 <= sh < 32 */
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_l(lo, hi, sh);
#else
	return (uint32_t)(((((uint64_t)hi << 32) | lo) << sh) >> 32);
#endif
}

/* k >>= w (0 < w < 32) */
template <int N> ECC_HD void shift_right(Fe<N> &k, int w)
{
#pragma unroll
	for (int j = 0; j < N - 1; j++) k.w[j] = funnel_r(k.w[j], k.w[j + 1], w);
	k.w[N - 1] >>= w;
}

/*
 * Fixed-base comb: acc = sum_i T[i][digit_i(k)],  T[i][d] = d * 2^(w*i) * G  (affine, Montgomery form, entry
 * (i << w) + d; d == 0 unused), 4 <= w <= 26.  One mixed addition per non-zero window, no doublings.  k must be < q.
 * For k < q the accumulator before window i is (k mod 2^(w*i))*G with 0 <= k mod 2^(w*i) < 2^(w*i) <= d*2^(w*i) < q,
 * so the add never meets P = +-Q; add_mixed resolves those cases anyway.
 */
template <class C> ECC_HD void load_table_entry(Aff<C> &t, const uint32_t *__restrict__ table, size_t e)
{
	constexpr int N = C::N;
	const uint32_t *base = table + e * (2 * N);
#if defined(__CUDA_ARCH__)
	if (N % 4 == 0) {
		const uint4 *src = reinterpret_cast<const uint4 *>(base);
#pragma unroll
		for (int j = 0; j < N / 4; j++) {
			uint4 v = __ldg(src + j);
			t.x.w[4 * j] = v.x;
			t.x.w[4 * j + 1] = v.y;
			t.x.w[4 * j + 2] = v.z;
			t.x.w[4 * j + 3] = v.w;
		}
#pragma unroll
		for (int j = 0; j < N / 4; j++) {
			uint4 v = __ldg(src + N / 4 + j);
			t.y.w[4 * j] = v.x;
			t.y.w[4 * j + 1] = v.y;
			t.y.w[4 * j + 2] = v.z;
			t.y.w[4 * j + 3] = v.w;
		}
	} else { /* N even: 8-byte loads */
		const uint2 *src = reinterpret_cast<const uint2 *>(base);
#pragma unroll
		for (int j = 0; j < N / 2; j++) {
			uint2 v = __ldg(src + j);
			t.x.w[2 * j] = v.x;
			t.x.w[2 * j + 1] = v.y;
		}
#pragma unroll
		for (int j = 0; j < N / 2; j++) {
			uint2 v = __ldg(src + N / 2 + j);
			t.y.w[2 * j] = v.x;
			t.y.w[2 * j + 1] = v.y;
		}
	}
#else
	for (int j = 0; j < N; j++) {
		t.x.w[j] = base[j];
		t.y.w[j] = base[N + j];
	}
#endif
}

template <class C> ECC_HD void comb_mul(Jac<C> &out, const Fe<C::N> &k, const uint32_t *__restrict__ table, int w)
{
	typedef EC<C> G;
	constexpr int N = C::N;
	typename G::XZ acc; /* extended Jacobian accumulator: 8M + 2S per window (xz_add_mixed) */
	G::xz_set_inf(acc);
	const int nwin = (C::QBITS + w - 1) / w;
	const uint32_t mask = (1u << w) - 1u;
	Fe<N> kk = k; /* consumed w bits at a time from the least significant end */
#pragma unroll 1
	for (int i = 0; i < nwin; i++) {
		uint32_t d = kk.w[0] & mask;
		shift_right<N>(kk, w);
#if defined(__CUDA_ARCH__) && defined(ECC_COMB_PREFETCH)
		/* the NEXT window's entry is requested into L2 while this window's addition runs: the gathers are random
		 * accesses into a table of tens of GB, i.e. DRAM latency on the critical path of every window otherwise */
		if (i + 1 < nwin) {
			const uint32_t nd = kk.w[0] & mask;
			const uint32_t *nxt = table + (((size_t)(i + 1) << w) + nd) * (2 * N);
			asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
			if ((2 * N * 4) % 128 != 0 && (2 * N * 4) > 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 2 * N - 1));
		}
#endif
		if (d != 0) {
			Aff<C> t;
			load_table_entry<C>(t, table, ((size_t)i << w) + d);
			typename G::XZ r;
			G::xz_add_mixed(r, acc, t);
			acc = r;
		}
	}
	G::xz_to_jac(out, acc); /* (X ZZ, Y ZZZ, ZZ): the Jacobian form K4 / the verification tail expect */
}

/* One field inversion for the calling thread alone: the inverter of the host build of the tests and of the one-off
 * table construction.  K2 / K3 pass an inverter that shares ONE inversion among the 128 threads of the CTA
 * (cta_inverse_128, kernels.cuh); an inverter of that kind must be called by every thread of the CTA. */
template <class C> struct ThreadInverter {
	ECC_HD void operator()(Fe<C::N> &r, const Fe<C::N> &a) const { Field<typename C::Fp>::inv(r, a); }
};

/*
 * Variable base: acc = k*P (+ addend), P affine and on the curve, k < q.  Signed 4-bit fixed window:
 * K' = k + 0x88..8 (one 8 per nibble); digit_i = nibble_i(K') - 8 in [-8, 7], plus a top digit = the carry out.
 * Table tbl[j] = (j+1)*P, j = 0..7, built with 7 mixed additions and then made AFFINE with one inversion
 * (Montgomery's trick over the seven Z's, and the addend's): every window addition is a mixed one (8M + 3S instead of
 * 12M + 4S), for ~50 products of conversion plus the thread's share of the inverter.  None of the multiples is the
 * point at infinity (the group order is a prime > 8).  The optional addend (uG of an ECDSA verification,
 * sig/ecdsa_common.c:796) is converted by the same inversion and added by one extra trip through the loop body, so the
 * kernel holds a single inlined copy of the addition.
 */
template <class C, class Inv>
ECC_HD void window_mul(Jac<C> &acc, const Fe<C::N> &k, const Aff<C> &P, const Jac<C> *addend, const Inv &invert)
{
	typedef EC<C> G;
	typedef Field<typename C::Fp> F;
	constexpr int N = C::N;
	Jac<C> tbl[8]; /* after the conversion only X, Y are meaningful (affine) */
	G::from_affine(tbl[0], P);
#pragma unroll 1
	for (int j = 1; j < 8; j++) G::add_mixed(tbl[j], tbl[j - 1], P); /* j == 1 takes the P == Q (doubling) branch */

	/* simultaneous inversion of Z(2P) .. Z(8P) and of the addend's Z */
	const bool have_add = addend != nullptr && !G::is_inf(*addend);
	Aff<C> add_aff;
	{
		Fe<N> pre[7]; /* pre[j] = Z(2P) * ... * Z((j+2)P) */
		pre[0] = tbl[1].Z;
#pragma unroll 1
		for (int j = 2; j < 8; j++) F::mul(pre[j - 1], pre[j - 2], tbl[j].Z);
		Fe<N> tot = pre[6], inv, zi, zi2, zi3, t;
		if (have_add) F::mul(tot, pre[6], addend->Z);
		invert(inv, tot);
		if (have_add) {
			F::mul(zi, inv, pre[6]);
			F::mul(t, inv, addend->Z);
			inv = t;
			F::sqr(zi2, zi);
			F::mul(zi3, zi2, zi);
			F::mul(add_aff.x, addend->X, zi2);
			F::mul(add_aff.y, addend->Y, zi3);
		}
#pragma unroll 1
		for (int j = 7; j >= 1; j--) {
			if (j > 1) {
				F::mul(zi, inv, pre[j - 2]);
				F::mul(t, inv, tbl[j].Z);
				inv = t;
			} else {
				zi = inv;
			}
			F::sqr(zi2, zi);
			F::mul(zi3, zi2, zi);
			F::mul(t, tbl[j].X, zi2);
			tbl[j].X = t;
			F::mul(t, tbl[j].Y, zi3);
			tbl[j].Y = t;
		}
	}

	/* K' = k + 0x88..8 over the ND = ceil(bitlen(q)/4) nibbles a reduced scalar occupies; the carry lands in
	 * nibble ND (0 or 1).  K' is then moved to the top of the N words so that nibbles leave from the MSB end. */
	constexpr int ND = (C::QBITS + 3) / 4;
	constexpr int PAD = 32 * N - 4 * ND; /* unused high bits: 0 for 256/384-bit, 52 for the 521-bit curve */
	uint32_t kk[N];
	uint64_t c = 0;
#pragma unroll
	for (int i = 0; i < N; i++) {
		const int nib = ND - 8 * i; /* nibbles of this word below ND */
		const uint32_t eights = nib >= 8 ? 0x88888888u : (nib <= 0 ? 0u : (0x88888888u >> (4 * (8 - nib))));
		uint64_t s = (uint64_t)k.w[i] + eights + c;
		kk[i] = (uint32_t)s;
		c = s >> 32;
	}
	if (PAD > 0) {
		constexpr int WS = PAD / 32, BS = PAD % 32;
		constexpr int TW = (4 * ND) / 32 < N ? (4 * ND) / 32 : 0; /* (index clamp only for the dead PAD == 0 case) */
		c = (kk[TW] >> ((4 * ND) % 32)) & 1u; /* top digit */
#pragma unroll
		for (int j = N - 1; j >= 0; j--) {
			uint32_t hi = (j - WS >= 0) ? kk[j - WS] : 0u;
			uint32_t lo = (j - WS - 1 >= 0) ? kk[j - WS - 1] : 0u;
			kk[j] = BS ? funnel_l(lo, hi, BS) : hi;
		}
	}
	G::set_inf(acc);
	if (c) G::from_affine(acc, P); /* top digit (weight 16^ND) is 0 or 1 */
#pragma unroll 1
	for (int di = ND - 1; di >= -1; di--) {
		Aff<C> e;
		Jac<C> t;
		bool have;
		if (di >= 0) {
#pragma unroll 1
			for (int q = 0; q < 4; q++) G::dbl(acc, acc);
			int d = (int)(kk[N - 1] >> 28) - 8; /* most significant nibble, then K' <<= 4 */
#pragma unroll
			for (int j = N - 1; j > 0; j--) kk[j] = funnel_l(kk[j - 1], kk[j], 4);
			kk[0] <<= 4;
			int ad = d < 0 ? -d : d;
			have = d != 0;
			const Jac<C> &te = tbl[(ad - 1) & 7];
			e.x = te.X;
			e.y = te.Y;
			if (d < 0) F::neg(e.y, e.y);
		} else {
			have = have_add;
			if (have) e = add_aff;
		}
		if (have) {
			G::add_mixed(t, acc, e);
			acc = t;
		}
	}
}

template <class C>
ECC_HD void window_mul(Jac<C> &acc, const Fe<C::N> &k, const Aff<C> &P, const Jac<C> *addend = nullptr)
{
	window_mul<C>(acc, k, P, addend, ThreadInverter<C>());
}

/* ---------------------------------------------------------------------------------------------- ECDSA */

/*
 * e = leftmost min(8*hlen, bitlen(q)) bits of the digest as an integer, reduced mod q:
 * steps 3-4 of __ecdsa_verify_finalize (sig/ecdsa_common.c:760-777).  Byte loads: hlen is arbitrary.
 */
template <class C> ECC_HD void digest_to_scalar(Fe<C::N> &e, const uint8_t *h, uint32_t hlen)
{
	constexpr int N = C::N;
	uint32_t qbytes = (C::QBITS + 7) / 8;
	uint32_t take = hlen < qbytes ? hlen : qbytes;
#pragma unroll
	for (int i = 0; i < N; i++) e.w[i] = 0;
	for (uint32_t i = 0; i < take; i++) {
		uint32_t pos = take - 1 - i; /* byte significance */
		uint32_t v = (uint32_t)h[i] << (8 * (pos & 3));
#pragma unroll
		for (int j = 0; j < N; j++) e.w[j] |= (j == (int)(pos >> 2)) ? v : 0u;
	}
	int sh = (int)(8 * take) - C::QBITS; /* > 0 only when bitlen(q) is not a multiple of 8 */
	if (sh > 0) {
#pragma unroll
		for (int j = 0; j < N; j++) {
			uint32_t hi = (j + 1 < N) ? e.w[j + 1] : 0u;
			e.w[j] = (e.w[j] >> sh) | (hi << (32 - sh));
		}
	}
	scalar_reduce<C>(e);
}

/* u = e * s^-1 mod q, v = r * s^-1 mod q (plain form): the mod-q scalar preparation of __ecdsa_verify_finalize
 * (sig/ecdsa_common.c:781-791: nn_modinv, nn_mod_mul), with s^-1 by Field::inv in the Montgomery domain of q. */
template <class C>
ECC_HD void ecdsa_uv(Fe<C::N> &u, Fe<C::N> &v, const Fe<C::N> &r, const Fe<C::N> &s, const Fe<C::N> &e)
{
	typedef Field<typename C::Fq> Fq;
	Fe<C::N> sm, wm;
	Fq::to_mont(sm, s);
	Fq::inv(wm, sm);   /* s^-1 * R mod q */
	Fq::mul(u, e, wm); /* (:786) */
	Fq::mul(v, r, wm); /* (:791) */
}

/*
 * ECDSA verification of one signature (r, s) on the reduced digest e under the public key Y (affine, validated,
 * Montgomery form).  Returns 0 = valid; 1 = r/s out of range, 2 = W' at infinity, 3 = r' != r (all map to -1).  Follows __ecdsa_verify_init's range checks (sig/ecdsa_common.c:653-658) and
 * __ecdsa_verify_finalize steps 5-10 (:781-810); differences that do not change the verdict:
 *   - s^-1 mod q by Field::inv (safegcd) in the Montgomery domain of q instead of nn_modinv's xgcd (:781);
 *   - W' = uG + vY stays Jacobian and "x(W') mod q == r" is tested without an inversion as X == c * Z^2 for the
 *     candidates c in {r, r+q} that are < p (:803-810);
 *   - uG through the comb table (K1), vY through the signed window (K2) instead of two ladders (:788,793).
 */
/* Steps 7-10 of __ecdsa_verify_finalize (sig/ecdsa_common.c:796-810) given u and v: W' = uG + vY, reject infinity,
 * accept iff x(W') mod q == r.  Returns 0 valid, 2 infinity, 3 mismatch. */
template <class C, class Inv>
ECC_HD int ecdsa_verify_tail(const Fe<C::N> &r, const Fe<C::N> &u, const Fe<C::N> &v_in, const Aff<C> &Y_in,
			     const uint32_t *__restrict__ table, int w, bool y_inf, const Inv &invert)
{
	typedef Field<typename C::Fp> F;
	constexpr int N = C::N;
	Jac<C> uG, W;
	comb_mul<C>(uG, u, table, w);
	/* A public key imported as the point at infinity (possible through the projective key formats,
	 * sig/ec_key.c:139): v*Y = infinity, exactly what the reference's complete formulas compute, so W' = uG.  The
	 * thread still walks the same code (v = 0 on a dummy base) because the inverter may be a CTA-wide one. */
	Fe<N> v = v_in;
	Aff<C> Y = Y_in;
	if (y_inf) {
#pragma unroll
		for (int i = 0; i < N; i++) {
			v.w[i] = 0;
			Y.x.w[i] = C::GX_MONT(i);
			Y.y.w[i] = C::GY_MONT(i);
		}
	}
	window_mul<C>(W, v, Y, &uG, invert); /* W' = vY + uG (:796) */
	if (EC<C>::is_inf(W)) return 2; /* (:799-800) */

	Fe<N> z2, c, t;
	F::sqr(z2, W.Z);
	bool match = false;
	if (!F::geq_mod(r)) { /* candidate x = r */
		F::to_mont(c, r);
		F::mul(t, c, z2);
		match = F::eq(t, W.X);
	}
	{ /* candidate x = r + q, when it is still a field element */
		Fe<N> rq;
		uint64_t cy = 0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t sum = (uint64_t)r.w[i] + C::Fq::P(i) + cy;
			rq.w[i] = (uint32_t)sum;
			cy = sum >> 32;
		}
		if (cy == 0 && !F::geq_mod(rq)) {
			F::to_mont(c, rq);
			F::mul(t, c, z2);
			match = match || F::eq(t, W.X);
		}
	}
	return match ? 0 : 3;
}

template <class C>
ECC_HD int ecdsa_verify_tail(const Fe<C::N> &r, const Fe<C::N> &u, const Fe<C::N> &v, const Aff<C> &Y,
			     const uint32_t *__restrict__ table, int w, bool y_inf = false)
{
	return ecdsa_verify_tail<C>(r, u, v, Y, table, w, y_inf, ThreadInverter<C>());
}

/* ------------------------------------------------------------------------------------------ ECFSDSA (§8f.4) */

/*
 * h = OS2I(digest) mod q with the WHOLE digest (sig/ecfsdsa.c:590-592: nn_init_from_buf + nn_mod — no truncation to
 * bitlen(q), unlike ECDSA), any hlen.  Horner over chunks of 4N bytes, most significant first:
 * h <- h * 2^(32N) + chunk (mod q).  A chunk (any integer < R) is reduced by a round trip through the Montgomery
 * domain (x -> xR -> x mod q), and h * 2^(32N) mod q = h * R mod q is exactly to_mont(h).
 */
template <class C> ECC_HD void digest_full_mod_q(Fe<C::N> &e, const uint8_t *h, uint32_t hlen)
{
	typedef Field<typename C::Fq> Fq;
	constexpr int N = C::N;
	const uint32_t nchunks = (hlen + 4u * N - 1u) / (4u * N);
#pragma unroll
	for (int i = 0; i < N; i++) e.w[i] = 0;
	for (uint32_t c = nchunks; c-- > 0;) {
		Fe<N> ch, t;
#pragma unroll
		for (int i = 0; i < N; i++) ch.w[i] = 0;
		const uint32_t lo = c * 4u * N, hi = (lo + 4u * N < hlen) ? lo + 4u * N : hlen; /* byte significances */
		for (uint32_t pos = lo; pos < hi; pos++) {
			uint32_t v = (uint32_t)h[hlen - 1 - pos] << (8 * (pos & 3));
			int wi = (int)((pos - lo) >> 2);
#pragma unroll
			for (int j = 0; j < N; j++) ch.w[j] |= (j == wi) ? v : 0u;
		}
		if (c + 1 < nchunks) {
			Fq::to_mont(t, e);
			e = t;
		}
		Fq::to_mont(t, ch);
		Fq::from_mont(ch, t);
		Fq::add(e, e, ch);
	}
}

/* Steps 5-7 of _ecfsdsa_verify_finalize (sig/ecfsdsa.c:597-610): W' = sG + eY with e = -h mod q, reject infinity
 * (prj_pt_unique fails on it), accept iff W' == r.  R is the signature's point (validated, Montgomery form); the
 * comparison is done projectively (X == r_x Z^2, Y == r_y Z^3) instead of normalising W'.  0 valid, 2 infinity,
 * 3 mismatch. */
template <class C, class Inv>
ECC_HD int ecfsdsa_verify_tail(const Aff<C> &R, const Fe<C::N> &s, const Fe<C::N> &e_neg, const Aff<C> &Y,
			       const uint32_t *__restrict__ table, int w, const Inv &invert)
{
	typedef Field<typename C::Fp> F;
	Jac<C> sG, W;
	comb_mul<C>(sG, s, table, w);
	window_mul<C>(W, e_neg, Y, &sG, invert);
	if (EC<C>::is_inf(W)) return 2;
	Fe<C::N> z2, z3, t;
	F::sqr(z2, W.Z);
	F::mul(z3, z2, W.Z);
	F::mul(t, R.x, z2);
	bool ok = F::eq(t, W.X);
	F::mul(t, R.y, z3);
	ok = ok && F::eq(t, W.Y);
	return ok ? 0 : 3;
}

/* ------------------------------------------------------------------------------------------ BIP0340 (§8f.4) */

/*
 * _bip0340_verify_finalize (sig/bip0340.c:497-577) after the init checks: W' = sG + (-e)Y' with Y' the public key
 * lifted to an even y (:540-545; the caller passes Y already lifted, Montgomery form), reject infinity (:555-556),
 * reject an odd y(W') (:559-560), accept iff x(W') == r (:563-564).  Unlike ECDSA / ECFSDSA the comparison needs the
 * affine representative (the parity of y), so W' is normalised with one more inversion — `invert`, a CTA-wide one in
 * the kernel, which every thread must call: rejected items walk the same code on dummy values.
 * r: plain integer < p.  0 valid, 2 infinity, 3 mismatch / odd y.
 */
template <class C, class Inv>
ECC_HD int bip0340_verify_tail(const Fe<C::N> &r, const Fe<C::N> &s, const Fe<C::N> &e_neg, const Aff<C> &Y,
			       const uint32_t *__restrict__ table, int w, const Inv &invert)
{
	typedef Field<typename C::Fp> F;
	Jac<C> sG, W;
	comb_mul<C>(sG, s, table, w);                   /* s may be 0: the comb then returns infinity (:537) */
	window_mul<C>(W, e_neg, Y, &sG, invert);
	const bool inf = EC<C>::is_inf(W);
	Fe<C::N> z, zi, zi2, zi3, x, y, t;
	z = W.Z;
	if (inf) F::set_one(z);
	invert(zi, z);
	F::sqr(zi2, zi);
	F::mul(zi3, zi2, zi);
	F::mul(t, W.X, zi2);
	F::from_mont(x, t);
	F::mul(t, W.Y, zi3);
	F::from_mont(y, t);
	if (inf) return 2;
	if (y.w[0] & 1u) return 3;
	return F::eq(x, r) ? 0 : 3;
}

/* r, s in [1, q-1]?  (__ecdsa_verify_init, sig/ecdsa_common.c:653-658) */
template <class C> ECC_HD bool ecdsa_rs_in_range(const Fe<C::N> &r, const Fe<C::N> &s)
{
	typedef Field<typename C::Fq> Fq;
	return !(Fq::is_zero(r) || Fq::is_zero(s) || Fq::geq_mod(r) || Fq::geq_mod(s));
}

/* Whole verification of one signature with a per-item inversion of s (host build of the tests and reference for the
 * kernel, which replaces the inversion by a CTA-wide simultaneous one). */
template <class C>
ECC_HD int ecdsa_verify_core(const Fe<C::N> &r, const Fe<C::N> &s, const Fe<C::N> &e, const Aff<C> &Y,
			     const uint32_t *__restrict__ table, int w)
{
	if (!ecdsa_rs_in_range<C>(r, s)) return 1;
	Fe<C::N> u, v;
	ecdsa_uv<C>(u, v, r, s, e);
	return ecdsa_verify_tail<C>(r, u, v, Y, table, w);
}

} // namespace eccb200
