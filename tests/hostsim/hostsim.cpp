/*
 * tests/hostsim/hostsim.cpp — TEST-ONLY host build of the device arithmetic (fp.cuh / ec.cuh compiled by g++).
 *
 * Purpose: let the CPU test-suite (`pytest -m "not gpu"`) check the ALGORITHMS the kernels use — word-level
 * Montgomery product, Jacobian formulas and their exceptional branches, comb / signed-window scalar
 * multiplication, the ECDSA verification core — against the oracle and the golden vectors without a GPU, and
 * count the field multiplications the implemented algorithm performs (M_impl).
 *
 * This is NOT a CPU fallback: it is built into tests/hostsim/_build/libecc_hostsim.so by the tests only, it is
 * never loaded by libecc_b200/, and the product library fails with -1 when no B200 is present.
 */
#define ECC_COUNT_MULS
#include "../../libecc_b200/csrc/ec.cuh"
#include "../../libecc_b200/csrc/msm_core.cuh"
#include "../../libecc_b200/csrc/sha3.cuh"
#include <cstring>
#include <vector>

namespace eccb200 {
thread_local unsigned long long g_fe_mul_count = 0;
}
using namespace eccb200;

template <class Fn> static int dispatch(int curve_id, Fn &&fn)
{
	switch (curve_id) {
	case 4: return fn(Curve_SECP256R1());
	case 1: return fn(Curve_FRP256V1());
	case 5: return fn(Curve_SECP384R1());
	case 8: return fn(Curve_BRAINPOOLP256R1());
	case 12: return fn(Curve_BRAINPOOLP384R1());
	case 19: return fn(Curve_SECP256K1());
	case 6: return fn(Curve_SECP521R1());
	case 17: return fn(Curve_SM2P256V1());
	case 9: return fn(Curve_BRAINPOOLP512R1());
	case 3: return fn(Curve_SECP224R1());
	case 2: return fn(Curve_SECP192R1());
	default: return -1;
	}
}

/* Jacobian -> affine big-endian bytes with a per-item inversion (the batched variant is kernel-only) */
template <class C> static int jac_to_wire(const Jac<C> &p, uint8_t *out)
{
	typedef Field<typename C::Fp> F;
	constexpr int N = C::N;
	memset(out, 0, 2 * C::PLEN);
	if (F::is_zero(p.Z)) return 1;
	Fe<N> zi, zi2, zi3, x, y, t;
	F::inv(zi, p.Z);
	F::sqr(zi2, zi);
	F::mul(zi3, zi2, zi);
	F::mul(t, p.X, zi2);
	F::from_mont(x, t);
	F::mul(t, p.Y, zi3);
	F::from_mont(y, t);
	store_be<N>(out, x, C::PLEN);
	store_be<N>(out + C::PLEN, y, C::PLEN);
	return 0;
}

template <class C> static bool load_point(Aff<C> &P, const uint8_t *buf)
{
	typedef Field<typename C::Fp> F;
	Fe<C::N> x, y;
	load_be<C::N>(x, buf, C::PLEN);
	load_be<C::N>(y, buf + C::PLEN, C::PLEN);
	bool ok = !F::geq_mod(x) && !F::geq_mod(y);
	F::to_mont(P.x, x);
	F::to_mont(P.y, y);
	return ok && EC<C>::on_curve(P);
}

/* comb table built with the same window_mul the device table kernel uses (small w only: this is a CPU) */
template <class C> static std::vector<uint32_t> build_table(int w)
{
	typedef Field<typename C::Fp> F;
	constexpr int N = C::N;
	int nwin = (C::QBITS + w - 1) / w;
	std::vector<uint32_t> tab((size_t)(nwin << w) * 2 * N, 0);
	Aff<C> G;
	for (int j = 0; j < N; j++) {
		G.x.w[j] = C::GX_MONT(j);
		G.y.w[j] = C::GY_MONT(j);
	}
	for (int i = 0; i < nwin; i++) {
		for (uint32_t d = 1; d < (1u << w); d++) {
			int bit = i * w;
			if (bit >= 32 * N) continue;
			Fe<N> k;
			for (int j = 0; j < N; j++) k.w[j] = 0;
			uint64_t v = (uint64_t)d << (bit & 31);
			k.w[bit >> 5] = (uint32_t)v;
			if ((bit >> 5) + 1 < N) k.w[(bit >> 5) + 1] = (uint32_t)(v >> 32);
			else if ((uint32_t)(v >> 32)) continue;
			if (Field<typename C::Fq>::geq_mod(k)) continue;
			Jac<C> acc;
			window_mul<C>(acc, k, G);
			Fe<N> zi, zi2, zi3, x, y;
			F::inv(zi, acc.Z);
			F::sqr(zi2, zi);
			F::mul(zi3, zi2, zi);
			F::mul(x, acc.X, zi2);
			F::mul(y, acc.Y, zi3);
			uint32_t *e = &tab[((size_t)(i << w) + d) * 2 * N];
			for (int j = 0; j < N; j++) {
				e[j] = x.w[j];
				e[N + j] = y.w[j];
			}
		}
	}
	return tab;
}

template <class C> static const std::vector<uint32_t> &table_for(int w)
{
	static std::vector<uint32_t> tabs[17];
	if (tabs[w].empty()) tabs[w] = build_table<C>(w);
	return tabs[w];
}

extern "C" {

int hostsim_fp_mul(int curve_id, int which, uint32_t n, const uint8_t *a, const uint8_t *b, uint8_t *out)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		constexpr int N = C::N;
		for (uint32_t i = 0; i < n; i++) {
			Fe<N> x, y, z;
			const int len = which == 0 ? C::PLEN : C::QLEN;
			load_be<N>(x, a + (size_t)i * len, len);
			load_be<N>(y, b + (size_t)i * len, len);
			if (which == 0) Field<typename C::Fp>::mul(z, x, y);
			else Field<typename C::Fq>::mul(z, x, y);
			store_be<N>(out + (size_t)i * len, z, len);
		}
		return 0;
	});
}

/* (a op b) mod p on plain values: op 0 add, 1 sub, 2 inverse of a (through the Montgomery domain) */
int hostsim_fp_op(int curve_id, int op, uint32_t n, const uint8_t *a, const uint8_t *b, uint8_t *out)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		typedef Field<typename C::Fp> F;
		constexpr int N = C::N;
		for (uint32_t i = 0; i < n; i++) {
			Fe<N> x, y, z;
			load_be<N>(x, a + (size_t)i * C::PLEN, C::PLEN);
			load_be<N>(y, b + (size_t)i * C::PLEN, C::PLEN);
			if (op == 0) F::add(z, x, y);
			else if (op == 1) F::sub(z, x, y);
			else {
				Fe<N> xm, zi;
				F::to_mont(xm, x);
				F::inv(zi, xm);
				F::from_mont(z, zi);
			}
			store_be<N>(out + (size_t)i * C::PLEN, z, C::PLEN);
		}
		return 0;
	});
}

/* out = a^-1 mod p (which = 0) or mod q (1), plain integers in and out, through Field::inv (safegcd, fermat = 0) or
 * Field::inv_fermat (1) */
int hostsim_fp_inv(int curve_id, int which, int fermat, uint32_t n, const uint8_t *a, uint8_t *out)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		constexpr int N = C::N;
		auto run = [&](auto ftag) {
			typedef decltype(ftag) FT;
			typedef Field<FT> F;
			for (uint32_t i = 0; i < n; i++) {
				Fe<N> x, xm, zi, z;
				load_be<N>(x, a + (size_t)i * FT::BYTES, FT::BYTES);
				F::to_mont(xm, x);
				if (fermat) F::inv_fermat(zi, xm);
				else F::inv(zi, xm);
				F::from_mont(z, zi);
				store_be<N>(out + (size_t)i * FT::BYTES, z, FT::BYTES);
			}
		};
		if (which == 0) run(typename C::Fp());
		else run(typename C::Fq());
		return 0;
	});
}

/* same contract as eccb200_prj_pt_mul_batch; w = comb window used for the fixed-base path */
int hostsim_prj_pt_mul_batch(int curve_id, int w, uint32_t n, const uint8_t *scalars, const uint8_t *points,
			     uint8_t *out, int8_t *status)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		constexpr int N = C::N;
		for (uint32_t i = 0; i < n; i++) {
			Fe<N> k;
			load_be<N>(k, scalars + (size_t)i * C::QLEN, C::QLEN);
			scalar_reduce<C>(k);
			Jac<C> acc;
			uint8_t *o = out + (size_t)i * 2 * C::PLEN;
			if (points) {
				Aff<C> P;
				if (!load_point<C>(P, points + (size_t)i * 2 * C::PLEN)) {
					memset(o, 0, 2 * C::PLEN);
					status[i] = -1;
					continue;
				}
				g_fe_mul_count = 0;
				window_mul<C>(acc, k, P);
			} else {
				const std::vector<uint32_t> &tab = table_for<C>(w);
				g_fe_mul_count = 0;
				comb_mul<C>(acc, k, tab.data(), w);
			}
			status[i] = (int8_t)jac_to_wire<C>(acc, o);
		}
		return 0;
	});
}

int hostsim_ecdsa_verify_batch(int curve_id, int w, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
			       const uint8_t *digests, uint32_t hlen, int8_t *verdict)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		constexpr int N = C::N;
		const std::vector<uint32_t> &tab = table_for<C>(w);
		for (uint32_t i = 0; i < n; i++) {
			Fe<N> r, s, e;
			load_be<N>(r, sigs + (size_t)i * 2 * C::QLEN, C::QLEN);
			load_be<N>(s, sigs + (size_t)i * 2 * C::QLEN + C::QLEN, C::QLEN);
			Aff<C> Y;
			bool ok = load_point<C>(Y, pubkeys + (size_t)i * 2 * C::PLEN);
			digest_to_scalar<C>(e, digests + (size_t)i * hlen, hlen);
			g_fe_mul_count = 0;
			ok = ok && (ecdsa_verify_core<C>(r, s, e, Y, tab.data(), w) == 0);
			verdict[i] = ok ? 0 : -1;
		}
		return 0;
	});
}

/* ECFSDSA verification with the kernel's building blocks (digest_full_mod_q, ecfsdsa_verify_tail) */
int hostsim_ecfsdsa_verify_batch(int curve_id, int w, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				 const uint8_t *digests, uint32_t hlen, int8_t *verdict)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		typedef Field<typename C::Fq> Fq;
		constexpr int N = C::N;
		const std::vector<uint32_t> &tab = table_for<C>(w);
		for (uint32_t i = 0; i < n; i++) {
			const uint8_t *sg = sigs + (size_t)i * (2 * C::PLEN + C::QLEN);
			Aff<C> R, Y;
			Fe<N> s, h;
			bool ok = load_point<C>(R, sg);
			load_be<N>(s, sg + 2 * C::PLEN, C::QLEN);
			ok = ok && !Fq::is_zero(s) && !Fq::geq_mod(s);
			ok = ok && load_point<C>(Y, pubkeys + (size_t)i * 2 * C::PLEN);
			digest_full_mod_q<C>(h, digests + (size_t)i * hlen, hlen);
			Fq::neg(h, h);
			ok = ok && (ecfsdsa_verify_tail<C>(R, s, h, Y, tab.data(), w, ThreadInverter<C>()) == 0);
			verdict[i] = ok ? 0 : -1;
		}
		return 0;
	});
}

/* W = a*G + b*Y with the kernel's building blocks (comb + signed window + one normalisation), status like the kernel */
int hostsim_double_smul_batch(int curve_id, int w, uint32_t n, const uint8_t *ab, const uint8_t *pubkeys, uint8_t *out,
			      int8_t *status)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		constexpr int N = C::N;
		const std::vector<uint32_t> &tab = table_for<C>(w);
		for (uint32_t i = 0; i < n; i++) {
			Fe<N> a, b;
			Aff<C> Y;
			load_be<N>(a, ab + (size_t)i * 2 * C::QLEN, C::QLEN);
			load_be<N>(b, ab + (size_t)i * 2 * C::QLEN + C::QLEN, C::QLEN);
			scalar_reduce<C>(a);
			scalar_reduce<C>(b);
			memset(out + (size_t)i * 2 * C::PLEN, 0, 2 * C::PLEN);
			if (!load_point<C>(Y, pubkeys + (size_t)i * 2 * C::PLEN)) {
				status[i] = -1;
				continue;
			}
			Jac<C> aG, W;
			comb_mul<C>(aG, a, tab.data(), w);
			window_mul<C>(W, b, Y, &aG, ThreadInverter<C>());
			status[i] = (int8_t)jac_to_wire<C>(W, out + (size_t)i * 2 * C::PLEN);
		}
		return 0;
	});
}

/* BIP0340 verification with the kernel's building blocks (digest_full_mod_q, bip0340_verify_tail) */
int hostsim_bip0340_verify_batch(int curve_id, int w, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				 const uint8_t *digests, uint32_t hlen, int8_t *verdict)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		typedef Field<typename C::Fq> Fq;
		typedef Field<typename C::Fp> F;
		constexpr int N = C::N;
		const std::vector<uint32_t> &tab = table_for<C>(w);
		for (uint32_t i = 0; i < n; i++) {
			const uint8_t *sg = sigs + (size_t)i * (C::PLEN + C::QLEN);
			const uint8_t *pkb = pubkeys + (size_t)i * 2 * C::PLEN;
			Aff<C> Y;
			Fe<N> r, s, h, yraw;
			load_be<N>(r, sg, C::PLEN);
			load_be<N>(s, sg + C::PLEN, C::QLEN);
			bool ok = !F::geq_mod(r) && !Fq::geq_mod(s);
			load_be<N>(yraw, pkb + C::PLEN, C::PLEN);
			ok = load_point<C>(Y, pkb) && ok;
			if (yraw.w[0] & 1u) F::neg(Y.y, Y.y);
			digest_full_mod_q<C>(h, digests + (size_t)i * hlen, hlen);
			Fq::neg(h, h);
			if (!ok) {
				verdict[i] = -1;
				continue;
			}
			verdict[i] = bip0340_verify_tail<C>(r, s, h, Y, tab.data(), w, ThreadInverter<C>()) == 0 ? 0 : -1;
		}
		return 0;
	});
}

/* group-law unit test: out = P1 + P2 on affine wire points through add_full / add_mixed / xz_add_mixed (which = 0 / 1 / 3),
 * or 2*P1 through dbl (which = 2); infinity operands are encoded as all-zero wire points */
int hostsim_point_op(int curve_id, int which, const uint8_t *p1, const uint8_t *p2, uint8_t *out, int8_t *status)
{
	return dispatch(curve_id, [&](auto c) {
		typedef decltype(c) C;
		typedef EC<C> G;
		constexpr int N = C::N;
		auto is_zero_buf = [](const uint8_t *b, int len) {
			for (int i = 0; i < len; i++)
				if (b[i]) return false;
			return true;
		};
		Jac<C> A, B, R;
		Aff<C> a, b;
		bool a_inf = is_zero_buf(p1, 2 * C::PLEN), b_inf = is_zero_buf(p2, 2 * C::PLEN);
		if (a_inf) G::set_inf(A);
		else {
			if (!load_point<C>(a, p1)) return -1;
			G::from_affine(A, a);
			/* use a non-trivial Z so the projective paths are exercised: (X,Y,Z) -> (4X, 8Y, 2Z) */
			typedef Field<typename C::Fp> F;
			F::dbl(A.Z, A.Z);
			F::dbl(A.X, A.X);
			F::dbl(A.X, A.X);
			F::dbl(A.Y, A.Y);
			F::dbl(A.Y, A.Y);
			F::dbl(A.Y, A.Y);
		}
		if (b_inf) G::set_inf(B);
		else {
			if (!load_point<C>(b, p2)) return -1;
			G::from_affine(B, b);
		}
		if (which == 0) G::add_full(R, A, B);
		else if (which == 1) {
			if (b_inf) return -1;
			G::add_mixed(R, A, b);
		} else if (which == 3) { /* extended Jacobian accumulator of the comb: (X, Y, ZZ, ZZZ) + affine */
			if (b_inf) return -1;
			typedef Field<typename C::Fp> F;
			typename G::XZ P, S;
			P.X = A.X;
			P.Y = A.Y;
			F::sqr(P.ZZ, A.Z);
			F::mul(P.ZZZ, P.ZZ, A.Z);
			G::xz_add_mixed(S, P, b);
			G::xz_to_jac(R, S);
		} else G::dbl(R, A);
		*status = (int8_t)jac_to_wire<C>(R, out);
		return 0;
	});
}

/*
 * ECFSDSA batch verification as one multi-scalar multiplication: the stages of libecc_b200/csrc/msm.cuh run serially
 * with the same building blocks (msm_core.cuh) — prepare, counting sort by bucket, bucket accumulation with the XYZZ
 * mixed addition, range reduction, per-window sums, Horner.  all_valid like eccb200_ecfsdsa_verify_msm_batch.
 * stats (optional, 5 entries): mixed additions of the accumulation, buckets, windows, field products in total, the
 * fullest bucket.
 */
int hostsim_schnorr_msm(int scheme, int curve_id, int c, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
			const uint8_t *digests, uint32_t hlen, const uint8_t *seed, int *all_valid, unsigned long long *stats)
{
	return dispatch(curve_id, [&](auto cv) {
		typedef decltype(cv) C;
		typedef Field<typename C::Fp> F;
		typedef Field<typename C::Fq> Fq;
		typedef EC<C> G;
		constexpr int N = C::N;
		*all_valid = 0;
		if (n == 0 || c < 2 || c > 16) return n == 0 ? 0 : -1;
		if (scheme == 2 && !msm_lift_supported<C>()) return -1;
		g_fe_mul_count = 0;
		MsmKey key;
		for (int i = 0; i < 8; i++)
			key.k[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
				   ((uint32_t)seed[4 * i + 3] << 24);
		const int nwin = msm_windows(C::QBITS - 1, c);
		const uint32_t nb = 1u << (c - 1), total = (uint32_t)nwin * nb, ch = nb < 16u ? nb : 16u, per_window = nb / ch;
		const size_t npts = 2 * (size_t)n + 1;
		std::vector<uint32_t> pts(npts * 2 * N), scal(npts * N);
		bool bad = false;
		Fe<N> ssum;
		Fq::set_zero(ssum);
		/* prepare (k_msm_prepare / k_msm_ssum) */
		for (uint32_t i = 0; i < n; i++) {
			Aff<C> W, Y, negW, Yf;
			Fe<N> s, h, a, cY, t;
			bool w_ok, key_ok;
			const uint8_t *pkb = pubkeys + (size_t)i * 2 * C::PLEN;
			if (scheme == 2) {
				const uint8_t *sg = sigs + (size_t)i * (C::PLEN + C::QLEN);
				Fe<N> r, yraw;
				load_be<N>(r, sg, C::PLEN);
				load_be<N>(s, sg + C::PLEN, C::QLEN);
				w_ok = !F::geq_mod(r);
				if (!w_ok) F::set_zero(r);
				w_ok = msm_lift_x<C>(W, r) && w_ok;
				load_be<N>(yraw, pkb + C::PLEN, C::PLEN);
				key_ok = load_point<C>(Y, pkb);
				if (yraw.w[0] & 1u) F::neg(Y.y, Y.y);
			} else {
				const uint8_t *sg = sigs + (size_t)i * (2 * C::PLEN + C::QLEN);
				w_ok = load_point<C>(W, sg);
				load_be<N>(s, sg + 2 * C::PLEN, C::QLEN);
				key_ok = load_point<C>(Y, pkb);
			}
			const bool s_ok = !Fq::geq_mod(s);
			digest_full_mod_q<C>(h, digests + (size_t)i * hlen, hlen);
			Fq::neg(h, h);
			msm_coefficient<N>(a, key, i, c);
			msm_terms<C>(negW, Yf, cY, t, W, Y, s, h, a);
			if (!(w_ok && s_ok && key_ok)) {
				bad = true;
				Fq::set_zero(a);
				Fq::set_zero(cY);
				Fq::set_zero(t);
			}
			msm_st<N>(&pts[(size_t)i * 2 * N], negW.x);
			msm_st<N>(&pts[(size_t)i * 2 * N + N], negW.y);
			msm_st<N>(&scal[(size_t)i * N], a);
			msm_st<N>(&pts[((size_t)n + i) * 2 * N], Yf.x);
			msm_st<N>(&pts[((size_t)n + i) * 2 * N + N], Yf.y);
			msm_st<N>(&scal[((size_t)n + i) * N], cY);
			Fq::add(ssum, ssum, t);
		}
		{
			Fe<N> gx, gy;
			for (int j = 0; j < N; j++) {
				gx.w[j] = C::GX_MONT(j);
				gy.w[j] = C::GY_MONT(j);
			}
			if (msm_fold<C>(ssum)) F::neg(gy, gy);
			msm_st<N>(&pts[(size_t)2 * n * 2 * N], gx);
			msm_st<N>(&pts[(size_t)2 * n * 2 * N + N], gy);
		}
		msm_st<N>(&scal[(size_t)2 * n * N], ssum);
		/* counting sort (k_msm_hist / k_msm_scan / k_msm_scatter) */
		std::vector<uint32_t> count(total, 0), offs(total, 0), fill(total, 0);
		for (size_t j = 0; j < npts; j++)
			msm_digits(&scal[j * N], N, c, nwin, [&](int w, int d) { count[(size_t)w * nb + (uint32_t)((d < 0 ? -d : d) - 1)]++; });
		uint32_t run = 0;
		for (uint32_t g = 0; g < total; g++) {
			offs[g] = run;
			run += count[g];
		}
		std::vector<uint32_t> list(run ? run : 1);
		for (size_t j = 0; j < npts; j++)
			msm_digits(&scal[j * N], N, c, nwin, [&](int w, int d) {
				const size_t g = (size_t)w * nb + (uint32_t)((d < 0 ? -d : d) - 1);
				list[offs[g] + fill[g]++] = (uint32_t)j | (d < 0 ? 0x80000000u : 0u);
			});
		/* accumulate (k_msm_accumulate) */
		std::vector<uint32_t> buckets((size_t)total * 3 * N);
		for (uint32_t g = 0; g < total; g++) {
			typename G::XZ acc;
			G::xz_set_inf(acc);
			for (uint32_t k = 0; k < count[g]; k++) {
				const uint32_t e = list[offs[g] + k];
				const uint32_t *pp = &pts[(size_t)(e & 0x7fffffffu) * 2 * N];
				Aff<C> P;
				msm_ld<N>(P.x, pp);
				msm_ld<N>(P.y, pp + N);
				if (e >> 31) F::neg(P.y, P.y);
				G::xz_add_mixed(acc, acc, P);
			}
			Jac<C> r;
			G::xz_to_jac(r, acc);
			msm_st_jac<C>(buckets.data(), g, r);
		}
		/* reduce (k_msm_reduce / k_msm_window_sum / k_msm_final) */
		std::vector<uint32_t> winsum((size_t)nwin * 3 * N);
		for (int w = 0; w < nwin; w++) {
			Jac<C> acc, part;
			G::set_inf(acc);
			for (uint32_t t = 0; t < per_window; t++) {
				msm_reduce_range<C>(part, buckets.data(), (size_t)w * nb, t * ch, ch);
				G::add_full(acc, acc, part);
			}
			msm_st_jac<C>(winsum.data(), (size_t)w, acc);
		}
		Jac<C> sum;
		msm_horner<C>(sum, winsum.data(), nwin, c);
		*all_valid = (G::is_inf(sum) && !bad) ? 1 : 0;
		if (stats) {
			stats[0] = run;
			stats[1] = total;
			stats[2] = (unsigned long long)nwin;
			stats[3] = g_fe_mul_count;
			unsigned long long mx = 0;
			for (uint32_t g = 0; g < total; g++) mx = count[g] > mx ? count[g] : mx;
			stats[4] = mx;
		}
		return 0;
	});
}

int hostsim_ecfsdsa_msm(int curve_id, int c, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys, const uint8_t *digests,
			uint32_t hlen, const uint8_t *seed, int *all_valid, unsigned long long *stats)
{
	return hostsim_schnorr_msm(1, curve_id, c, n, sigs, pubkeys, digests, hlen, seed, all_valid, stats);
}

int hostsim_bip0340_msm(int curve_id, int c, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys, const uint8_t *digests,
			uint32_t hlen, const uint8_t *seed, int *all_valid, unsigned long long *stats)
{
	return hostsim_schnorr_msm(2, curve_id, c, n, sigs, pubkeys, digests, hlen, seed, all_valid, stats);
}

/* square root modulo p = 3 mod 4 as the lift of r uses it: out = sqrt(a) (either root), returns 1 when a is a square, 0
 * when not, -1 for p = 1 mod 4.  a, out: PLEN big-endian bytes, plain integers. */
int hostsim_fp_sqrt(int curve_id, const uint8_t *a, uint8_t *out)
{
	return dispatch(curve_id, [&](auto cv) {
		typedef decltype(cv) C;
		typedef Field<typename C::Fp> F;
		if (!msm_lift_supported<C>()) return -1;
		Fe<C::N> x, xm, r, rp;
		load_be<C::N>(x, a, C::PLEN);
		F::to_mont(xm, x);
		const bool ok = F::sqrt_3mod4(r, xm);
		F::from_mont(rp, r);
		store_be<C::N>(out, rp, C::PLEN);
		return ok ? 1 : 0;
	});
}

/* signed digits of a scalar (little-endian 32-bit words), for the recoding test: digits[w] for w < nwin; returns nwin */
int hostsim_msm_digits(const uint32_t *k, int nwords, int bits, int c, int *digits)
{
	const int nwin = msm_windows(bits, c);
	for (int w = 0; w < nwin; w++) digits[w] = 0;
	msm_digits(k, nwords, c, nwin, [&](int w, int d) { digits[w] = d; });
	return nwin;
}

void hostsim_msm_coefficient(const uint8_t *seed, uint64_t i, uint32_t out[8])
{
	MsmKey key;
	for (int j = 0; j < 8; j++)
		key.k[j] = (uint32_t)seed[4 * j] | ((uint32_t)seed[4 * j + 1] << 8) | ((uint32_t)seed[4 * j + 2] << 16) |
			   ((uint32_t)seed[4 * j + 3] << 24);
	msm_chacha20_block8(out, key, i);
}

unsigned long long hostsim_last_mul_count(void) { return g_fe_mul_count; }

/* SHA-3 as compiled for the device (sha3.cuh is plain C++) */
void hostsim_sha3(int digest_bytes, const uint8_t *m, uint64_t len, uint8_t *out) { sha3_device(m, len, out, digest_bytes); }

} /* extern "C" */
