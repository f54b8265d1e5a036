/*
 * msm_core.cuh — building blocks of batch verification as ONE multi-scalar multiplication.
 *
 * The reference's Schnorr-type verify_batch implementations (sig/ecfsdsa.c:711-1074, sig/bip0340.c:1296) check a whole
 * batch with one random linear combination,
 *         (sum a_i s_i) G  +  sum (-a_i) W_i  +  sum (-a_i e_i) Y_i  ==  infinity,
 * evaluated with the Bos-Coster heap (sig/sig_algs.c:1052).  A heap is a serial structure; the data-parallel form of the
 * same sum is the bucket method (Pippenger): cut every scalar into signed c-bit digits, add each point into the bucket
 * of its digit (one mixed addition per point and window — no doublings), then reduce the buckets of a window with running
 * sums and combine the windows by Horner.  For 2^20 signatures and c = 16 that is ~26 mixed additions per signature
 * instead of the ~2 000 field products of an individual verification.
 *
 * This file holds the pieces that run unchanged on the device (msm.cuh wraps them in kernels) and on the host build of the
 * tests (tests/hostsim): the coefficient generator, the signed-digit recoding, the per-signature preparation, the bucket
 * range reduction and the final Horner combination.
 */
#pragma once
#include "ec.cuh"

namespace eccb200 {

/* ---------------------------------------------------------------------------------------------- coefficients a_i */

/*
 * a_i = the leading c * ceil(128 / c) - 1 bits (127 .. 139, see msm_coefficient_bits) of the ChaCha20 block (RFC 8439
 * block function) keyed by the per-call 256-bit seed, with block counter i.  The reference draws a_i with
 * nn_get_random_mod (sig/ecfsdsa.c:915); any coefficients the signer cannot predict give the same guarantee, and short
 * ones halve the work on the W_i (a batch holding a forgery passes with probability <= 2^-127).
 */
struct MsmKey {
	uint32_t k[8];
};

#define ECC_CHACHA_QR(a, b, c, d)             \
	do {                                  \
		a += b; d ^= a; d = (d << 16) | (d >> 16); \
		c += d; b ^= c; b = (b << 12) | (b >> 20); \
		a += b; d ^= a; d = (d << 8) | (d >> 24);  \
		c += d; b ^= c; b = (b << 7) | (b >> 25);  \
	} while (0)

ECC_HD void msm_chacha20_block8(uint32_t out[8], const MsmKey &key, uint64_t counter)
{
	const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
	const uint32_t n0 = (uint32_t)counter, n1 = (uint32_t)(counter >> 32), n2 = 0x314d534du /* "MSM1" */, n3 = 0;
	uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, x4 = key.k[0], x5 = key.k[1], x6 = key.k[2], x7 = key.k[3],
		 x8 = key.k[4], x9 = key.k[5], x10 = key.k[6], x11 = key.k[7], x12 = n0, x13 = n1, x14 = n2, x15 = n3;
	for (int r = 0; r < 10; r++) {
		ECC_CHACHA_QR(x0, x4, x8, x12);
		ECC_CHACHA_QR(x1, x5, x9, x13);
		ECC_CHACHA_QR(x2, x6, x10, x14);
		ECC_CHACHA_QR(x3, x7, x11, x15);
		ECC_CHACHA_QR(x0, x5, x10, x15);
		ECC_CHACHA_QR(x1, x6, x11, x12);
		ECC_CHACHA_QR(x2, x7, x8, x13);
		ECC_CHACHA_QR(x3, x4, x9, x14);
	}
	out[0] = x0 + c0;
	out[1] = x1 + c1;
	out[2] = x2 + c2;
	out[3] = x3 + c3;
	out[4] = x4 + key.k[0];
	out[5] = x5 + key.k[1];
	out[6] = x6 + key.k[2];
	out[7] = x7 + key.k[3];
}

/* ---------------------------------------------------------------------------------------------- signed digits */

/*
 * A scalar below 2^bits in signed base-2^c digits d_w in [-2^(c-1), 2^(c-1)] takes bits / c + 1 windows: the top
 * window holds the remaining t = bits mod c bits plus the carry, at most 2^t <= 2^(c-1), so nothing is carried out.
 * Two measures keep the top window from collecting points in a handful of buckets (one thread adds up one bucket):
 *   - the coefficients a_i have c * ceil(128 / c) - 1 bits (>= 127): their top window is a full one minus one bit;
 *   - the full-width scalars are folded to k <= (q - 1) / 2 by negating the point (msm_fold), so they have
 *     bitlen(q) - 1 bits and, for every byte-aligned order at c = 16, a 15-bit top window.
 */
ECC_HD int msm_windows(int bits, int c) { return bits / c + 1; }
ECC_HD int msm_coefficient_bits(int c) { return c * ((128 + c - 1) / c) - 1; }

template <int N> ECC_HD void msm_coefficient(Fe<N> &a, const MsmKey &key, uint64_t i, int c)
{
	uint32_t o[8];
	msm_chacha20_block8(o, key, i);
	const int bits = msm_coefficient_bits(c); /* 127 .. 139 */
#pragma unroll
	for (int j = 0; j < N; j++) {
		uint32_t v = j < 8 ? o[j < 8 ? j : 0] : 0u;
		const int lo = 32 * j;
		if (lo >= bits) v = 0;
		else if (lo + 32 > bits) v &= (1u << (bits - lo)) - 1u;
		a.w[j] = v;
	}
}

/* k <- min(k, q - k); returns true when the point has to be negated (k and q - k are both below q; k = 0 stays) */
template <class C> ECC_HD bool msm_fold(Fe<C::N> &k)
{
	typedef Field<typename C::Fq> Fq;
	Fe<C::N> kn;
	Fq::neg(kn, k); /* q - k, or 0 for k = 0 */
	uint64_t bw = 0; /* borrow of kn - k: set iff kn < k */
#pragma unroll
	for (int i = 0; i < C::N; i++) {
		const uint64_t t = (uint64_t)kn.w[i] - k.w[i] - bw;
		bw = (t >> 32) & 1;
	}
	const bool take = bw != 0 && !Fq::is_zero(k);
	if (take) k = kn;
	return take;
}

/* calls f(w, d) for every non-zero digit of the nwords-word scalar k (k read through a pointer: global memory on the
 * device, so the dynamic word index costs nothing); 2 <= c <= 16 */
template <class Fn> ECC_HD void msm_digits(const uint32_t *k, int nwords, int c, int nwin, Fn &&f)
{
	const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1u;
	uint32_t carry = 0;
	for (int w = 0; w < nwin; w++) {
		const int pos = w * c, wi = pos >> 5, off = pos & 31;
		const uint32_t lo = wi < nwords ? k[wi] : 0u, hi = wi + 1 < nwords ? k[wi + 1] : 0u;
		const uint32_t raw = ((off ? ((lo >> off) | (hi << (32 - off))) : lo) & mask) + carry;
		int d;
		if (raw > half) {
			d = (int)raw - (int)(mask + 1u);
			carry = 1;
		} else {
			d = (int)raw;
			carry = 0;
		}
		if (d) f(w, d);
	}
}

/* ---------------------------------------------------------------------------------------------- word buffers */

template <int N> ECC_HD void msm_ld(Fe<N> &r, const uint32_t *src)
{
#if defined(__CUDA_ARCH__)
	if (N % 4 == 0) {
		const uint4 *p = reinterpret_cast<const uint4 *>(src);
#pragma unroll
		for (int j = 0; j < N / 4; j++) {
			const uint4 v = p[j];
			r.w[4 * j] = v.x;
			r.w[4 * j + 1] = v.y;
			r.w[4 * j + 2] = v.z;
			r.w[4 * j + 3] = v.w;
		}
		return;
	}
#endif
#pragma unroll
	for (int j = 0; j < N; j++) r.w[j] = src[j];
}

template <int N> ECC_HD void msm_st(uint32_t *dst, const Fe<N> &a)
{
#if defined(__CUDA_ARCH__)
	if (N % 4 == 0) {
		uint4 *p = reinterpret_cast<uint4 *>(dst);
#pragma unroll
		for (int j = 0; j < N / 4; j++) {
			uint4 v;
			v.x = a.w[4 * j];
			v.y = a.w[4 * j + 1];
			v.z = a.w[4 * j + 2];
			v.w = a.w[4 * j + 3];
			p[j] = v;
		}
		return;
	}
#endif
#pragma unroll
	for (int j = 0; j < N; j++) dst[j] = a.w[j];
}

/* points travel between the stages as words in Montgomery form: rows of 2N (affine) or 3N (Jacobian) words, 16-byte
 * aligned whenever N is a multiple of 4 (the other curves take the word loop above) */
template <class C> ECC_HD void msm_ld_jac(Jac<C> &p, const uint32_t *buf, size_t idx)
{
	const uint32_t *b = buf + idx * (3 * C::N);
	msm_ld<C::N>(p.X, b);
	msm_ld<C::N>(p.Y, b + C::N);
	msm_ld<C::N>(p.Z, b + 2 * C::N);
}
template <class C> ECC_HD void msm_st_jac(uint32_t *buf, size_t idx, const Jac<C> &p)
{
	uint32_t *b = buf + idx * (3 * C::N);
	msm_st<C::N>(b, p.X);
	msm_st<C::N>(b + C::N, p.Y);
	msm_st<C::N>(b + 2 * C::N, p.Z);
}

/* ---------------------------------------------------------------------------------------------- per signature */

/*
 * ECFSDSA (sig/ecfsdsa.c:881-993): signature W_x || W_y || s, digest H(W_x || W_y || m).  Given the parsed, validated
 * pieces and the coefficient a, produce the two terms of the sum this signature owns,
 *       a * (-W)        (-W stored, so the short a is the scalar: the reference multiplies W by -a mod q, :985-987)
 *       (a * e) * Y     with e = -h mod q (:963-968), folded to a scalar <= (q - 1) / 2 on +-Y
 * and its share t = a * s mod q of the generator's scalar (:925-927).  All scalars in plain form.
 */
template <class C>
ECC_HD void msm_terms(Aff<C> &negW, Aff<C> &Yf, Fe<C::N> &cY, Fe<C::N> &t, const Aff<C> &W, const Aff<C> &Y,
		      const Fe<C::N> &s, const Fe<C::N> &e_neg, const Fe<C::N> &a)
{
	typedef Field<typename C::Fp> F;
	typedef Field<typename C::Fq> Fq;
	Fe<C::N> am;
	negW.x = W.x;
	F::neg(negW.y, W.y);
	Fq::to_mont(am, a);      /* a R mod q */
	Fq::mul(cY, am, e_neg);  /* a e mod q, plain */
	Fq::mul(t, am, s);       /* a s mod q, plain */
	Yf = Y;
	if (msm_fold<C>(cY)) F::neg(Yf.y, Y.y);
}

/*
 * BIP0340 (sig/bip0340.c:1166-1200): the signature carries only r = x(R); the batch form needs the point, so R is lifted
 * from r with the even y (aff_pt_y_from_x + the parity choice, :1188-1196).  r: plain integer < p.  Returns false when
 * r^3 + a r + b is not a square (the reference's fp_sqrt fails and the batch is rejected).  Primes p = 3 mod 4 only
 * (every curve of this library except SECP224R1): one exponentiation instead of Tonelli-Shanks.
 */
template <class C> ECC_HD bool msm_lift_x(Aff<C> &R, const Fe<C::N> &r)
{
	typedef Field<typename C::Fp> F;
	Fe<C::N> v, y, yp;
	F::to_mont(R.x, r);
	EC<C>::curve_rhs(v, R.x);
	const bool ok = F::sqrt_3mod4(y, v);
	F::from_mont(yp, y);
	if (yp.w[0] & 1u) F::neg(y, y);
	R.y = y;
	return ok;
}
template <class C> ECC_HD constexpr bool msm_lift_supported() { return (C::Fp::P(0) & 3u) == 3u; }

/* ---------------------------------------------------------------------------------------------- bucket reduction */

/*
 * Buckets lo .. lo+ch-1 of one window (bucket b holds the sum of the points whose digit is +-(b+1)):
 *       sum_{b} (b + 1) B_b  =  sum_{b} (b - lo + 1) B_b  +  lo * sum_{b} B_b
 * the first term by the running-sum trick (2 ch additions), the second by a short double-and-add on the range total.
 */
template <class C>
ECC_HD void msm_reduce_range(Jac<C> &out, const uint32_t *buckets, size_t first_bucket, uint32_t lo, uint32_t ch)
{
	typedef EC<C> G;
	Jac<C> run, tot, bk;
	G::set_inf(run);
	G::set_inf(tot);
	for (uint32_t b = ch; b-- > 0;) {
		msm_ld_jac<C>(bk, buckets, first_bucket + lo + b);
		G::add_full_ool(run, run, bk);
		G::add_full_ool(tot, tot, run);
	}
	if (lo) {
		Jac<C> m;
		G::set_inf(m);
		int top = 31;
		while (!((lo >> top) & 1u)) top--;
		for (int bit = top; bit >= 0; bit--) {
			G::dbl(m, m);
			if ((lo >> bit) & 1u) G::add_full_ool(m, m, run);
		}
		G::add_full_ool(tot, tot, m);
	}
	out = tot;
}

/* sum_w 2^(c w) S_w by Horner, most significant window first */
template <class C> ECC_HD void msm_horner(Jac<C> &acc, const uint32_t *winsum, int nwin, int c)
{
	typedef EC<C> G;
	Jac<C> s;
	msm_ld_jac<C>(acc, winsum, (size_t)(nwin - 1));
	for (int w = nwin - 2; w >= 0; w--) {
		for (int i = 0; i < c; i++) G::dbl(acc, acc);
		msm_ld_jac<C>(s, winsum, (size_t)w);
		G::add_full_ool(acc, acc, s);
	}
}

} // namespace eccb200
