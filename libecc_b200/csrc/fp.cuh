/*
 * fp.cuh — prime-field arithmetic for the device (replaces the reference's src/nn + src/fp on the hot path).
 *
 * Reference counterparts (paths relative to /root/reference/src):
 *   Field<F>::mul   fp_mul_monty  fp/fp_montgomery.c:44 -> nn_mul_redc1 nn/nn_mul_redc1.c:124-218 (CIOS)
 *   Field<F>::sqr   fp_sqr_monty  fp/fp_montgomery.c:53
 *   Field<F>::add   fp_add_monty  fp/fp_montgomery.c:26 -> nn_mod_add nn/nn_add.c:337
 *   Field<F>::sub   fp_sub_monty  fp/fp_montgomery.c:35 -> nn_mod_sub nn/nn_add.c:398
 *   Field<F>::inv   fp_inv        fp/fp_mul.c:51 -> nn_modinv_fermat_redc nn/nn_modinv.c:538
 *   to_mont/from_mont  fp_redcify / fp_unredcify  fp/fp_mul_redc1.c:62,79
 *
 * Representation: an element is N 32-bit little-endian words held in registers (N = 8 for 256-bit, 12 for
 * 384-bit), always fully reduced (< m), in Montgomery form with R = 2^(32N) — the same R as the reference's
 * 64-bit-limb R = 2^(64n), so Montgomery-form values are bit-identical to the reference's.
 * One thread owns whole elements; there is no cross-lane traffic in the arithmetic (DESIGN.md §3 records the
 * measurement behind that choice).
 *
 * The modulus is a compile-time constant of the field tag F (curve_constants.inc), so the compiler folds the
 * special words of the NIST primes (0, 1, 0xffffffff, M0 == 1) out of the reduction.
 *
 * Two multiplier back ends produce bit-identical results:
 *   - portable C++ (this file): CIOS over 32x32->64 products; also compiles for the host so that the unit tests
 *     in tests/hostsim can check the algorithms without a GPU (test-only; the product never runs on the host);
 *   - PTX (fp_ptx.cuh, device only): IMAD.WIDE carry chains, the default on the device (-DECC_NO_PTX selects the portable one).
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define ECC_HD __host__ __device__ __forceinline__
#define ECC_D __device__ __forceinline__
#define ECC_NOINLINE __host__ __device__ __noinline__
#else
#define ECC_HD inline
#define ECC_D inline
#define ECC_NOINLINE __attribute__((noinline))
#endif

#define ECC_CONST_ARRAY(NAME, CNT, ...)                      \
	static constexpr ECC_HD uint32_t NAME(int i)          \
	{                                                     \
		constexpr uint32_t v[CNT] = { __VA_ARGS__ }; \
		return v[i];                                  \
	}

namespace eccb200 {

#include "curve_constants.inc"

template <int N> struct Fe {
	uint32_t w[N];
};

#if defined(ECC_COUNT_MULS)
/* host-only instrumentation (tests/hostsim): counts field multiplications, the M_impl of SURVEY.md §8d */
extern thread_local unsigned long long g_fe_mul_count;
#define ECC_COUNT_MUL() (++g_fe_mul_count)
#else
#define ECC_COUNT_MUL() ((void)0)
#endif

template <class F> struct FieldPortable {
	static constexpr int N = F::N;
	typedef Fe<N> E;

	/* r = a*b*R^-1 mod m.  32-bit-word CIOS: same recurrence as nn_mul_redc1.c:175-205 with WORD_BITS = 32. */
	static ECC_HD void mul(E &r, const E &a, const E &b)
	{
		uint32_t t[N + 2];
		ECC_COUNT_MUL();
#pragma unroll
		for (int i = 0; i < N + 2; i++) t[i] = 0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t c = 0;
#pragma unroll
			for (int j = 0; j < N; j++) {
				uint64_t s = (uint64_t)a.w[j] * b.w[i] + t[j] + c;
				t[j] = (uint32_t)s;
				c = s >> 32;
			}
			uint64_t s = (uint64_t)t[N] + c;
			t[N] = (uint32_t)s;
			t[N + 1] = (uint32_t)(s >> 32);
			uint32_t m = t[0] * F::M0;
			c = ((uint64_t)m * F::P(0) + t[0]) >> 32;
#pragma unroll
			for (int j = 1; j < N; j++) {
				uint64_t s2 = (uint64_t)m * F::P(j) + t[j] + c;
				t[j - 1] = (uint32_t)s2;
				c = s2 >> 32;
			}
			s = (uint64_t)t[N] + c;
			t[N - 1] = (uint32_t)s;
			t[N] = t[N + 1] + (uint32_t)(s >> 32);
		}
		/* t < 2m: conditional subtraction (nn_mul_redc1.c:210-211) */
		uint32_t d[N];
		uint64_t bw = 0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t s = (uint64_t)t[i] - F::P(i) - bw;
			d[i] = (uint32_t)s;
			bw = (s >> 32) & 1;
		}
		bool ge = (t[N] != 0) || (bw == 0);
#pragma unroll
		for (int i = 0; i < N; i++) r.w[i] = ge ? d[i] : t[i];
	}

	static ECC_HD void sqr(E &r, const E &a) { mul(r, a, a); }

	static ECC_HD void add(E &r, const E &a, const E &b)
	{
		uint32_t s[N], d[N];
		uint64_t c = 0, bw = 0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t t = (uint64_t)a.w[i] + b.w[i] + c;
			s[i] = (uint32_t)t;
			c = t >> 32;
		}
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t t = (uint64_t)s[i] - F::P(i) - bw;
			d[i] = (uint32_t)t;
			bw = (t >> 32) & 1;
		}
		bool ge = (c != 0) || (bw == 0);
#pragma unroll
		for (int i = 0; i < N; i++) r.w[i] = ge ? d[i] : s[i];
	}

	static ECC_HD void sub(E &r, const E &a, const E &b)
	{
		uint32_t d[N];
		uint64_t bw = 0, c = 0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t t = (uint64_t)a.w[i] - b.w[i] - bw;
			d[i] = (uint32_t)t;
			bw = (t >> 32) & 1;
		}
		uint32_t mask = (uint32_t)0 - (uint32_t)bw;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t t = (uint64_t)d[i] + (F::P(i) & mask) + c;
			r.w[i] = (uint32_t)t;
			c = t >> 32;
		}
	}
};

} // namespace eccb200

#if defined(__CUDA_ARCH__) && !defined(ECC_NO_PTX)
#include "fp_ptx.cuh"
#endif

namespace eccb200 {

#if defined(__CUDA_ARCH__) && !defined(ECC_NO_PTX)
template <class F> struct FieldCore : FieldPtx<F> {};
#else
template <class F> struct FieldCore : FieldPortable<F> {};
#endif

template <class F> struct Field : FieldCore<F> {
	static constexpr int N = F::N;
	typedef Fe<N> E;
	typedef FieldCore<F> Core;
	using Core::add;
	using Core::mul;
	using Core::sqr;
	using Core::sub;

	static ECC_HD void set_zero(E &r)
	{
#pragma unroll
		for (int i = 0; i < N; i++) r.w[i] = 0;
	}
	static ECC_HD void set_one(E &r) /* 1 in Montgomery form */
	{
#pragma unroll
		for (int i = 0; i < N; i++) r.w[i] = F::ONE(i);
	}
	static ECC_HD bool is_zero(const E &a)
	{
		uint32_t acc = 0;
#pragma unroll
		for (int i = 0; i < N; i++) acc |= a.w[i];
		return acc == 0;
	}
	static ECC_HD bool eq(const E &a, const E &b)
	{
		uint32_t acc = 0;
#pragma unroll
		for (int i = 0; i < N; i++) acc |= a.w[i] ^ b.w[i];
		return acc == 0;
	}
	/* a >= m ? (raw integer comparison; used to validate wire inputs) */
	static ECC_HD bool geq_mod(const E &a)
	{
		uint64_t bw = 0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t t = (uint64_t)a.w[i] - F::P(i) - bw;
			bw = (t >> 32) & 1;
		}
		return bw == 0;
	}
	/* r = a - m if a >= m else a  (raw integers < 2^(32N)) */
	static ECC_HD void cond_sub_mod(E &r, const E &a)
	{
		uint32_t d[N];
		uint64_t bw = 0;
#pragma unroll
		for (int i = 0; i < N; i++) {
			uint64_t t = (uint64_t)a.w[i] - F::P(i) - bw;
			d[i] = (uint32_t)t;
			bw = (t >> 32) & 1;
		}
#pragma unroll
		for (int i = 0; i < N; i++) r.w[i] = bw ? a.w[i] : d[i];
	}
	static ECC_HD void neg(E &r, const E &a)
	{
		E z;
		set_zero(z);
		sub(r, z, a);
	}
	static ECC_HD void dbl(E &r, const E &a) { add(r, a, a); }
	static ECC_HD void to_mont(E &r, const E &a)
	{
		E rr;
#pragma unroll
		for (int i = 0; i < N; i++) rr.w[i] = F::RR(i);
		mul(r, a, rr);
	}
	static ECC_HD void from_mont(E &r, const E &a)
	{
		E one;
#pragma unroll
		for (int i = 0; i < N; i++) one.w[i] = (i == 0) ? 1u : 0u;
		mul(r, a, one);
	}
	static ECC_HD void cmov(E &r, const E &a, bool take)
	{
#pragma unroll
		for (int i = 0; i < N; i++) r.w[i] = take ? a.w[i] : r.w[i];
	}

	/*
	 * r = a^(m-2) (Montgomery in, Montgomery out): Fermat inversion, 4-bit fixed window over the constant exponent
	 * (fp_inv, fp/fp_mul.c:51 -> nn_mod_pow_redc, nn/nn_mod_pow.c:39 is a bit-by-bit ladder; same result).
	 * a == 0 gives 0.
	 */
	static ECC_HD void inv(E &r, const E &a)
	{
		E tbl[16];
		set_one(tbl[0]);
		tbl[1] = a;
#pragma unroll 1
		for (int i = 2; i < 16; i++) mul(tbl[i], tbl[i - 1], a);
		E acc;
		set_one(acc);
#pragma unroll 1
		for (int wi = N - 1; wi >= 0; wi--) {
			uint32_t ew = pm2_word(wi);
#pragma unroll 1
			for (int nb = 7; nb >= 0; nb--) {
#pragma unroll 1
				for (int q = 0; q < 4; q++) sqr(acc, acc);
				uint32_t d = (ew >> (4 * nb)) & 15u;
				mul(acc, acc, tbl[d]);
			}
		}
		r = acc;
	}

      private:
	static ECC_HD uint32_t pm2_word(int i)
	{
		uint32_t v = 0;
#pragma unroll
		for (int k = 0; k < N; k++) v = (k == i) ? F::PM2(k) : v;
		return v;
	}
};

} // namespace eccb200
