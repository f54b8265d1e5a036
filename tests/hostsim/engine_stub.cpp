/*
 * tests/hostsim/engine_stub.cpp — TEST-ONLY stand-in for libecc_b200.so: the engine entry points the drop-in layer
 * (libecc_b200/csrc/dropin.cpp) calls, served by the host build of the device algorithms (hostsim.cpp).
 *
 * Purpose: let `pytest -m "not gpu"` run the drop-in's HOST logic — struct marshalling, scheme checks, mod-q scalar
 * preparation, hashing through the reference's src/hash, verdict mapping, forwarding — through the C harness
 * (tests/dropin/dropin_harness.c) against the unmodified reference without a GPU: the harness is started with
 * LD_PRELOAD=tests/hostsim/_build/libecc_b200_stub.so, so the drop-in's eccb200_* references bind here instead of to the
 * product library.  This is not a CPU fallback: nothing under libecc_b200/ loads it, the product library still fails
 * with -1 without a B200, and the `-m gpu` tests run the same harness on the real engine.
 *
 * Contracts: include/libecc_b200.h (same arguments, status codes and edge behaviour as the entry points replaced).
 */
#include "hostsim.cpp"
#include <cstdlib>
#include <mutex>

struct eccb200_ctx {
	int curve_id;
	int w; /* comb window of the host table: small, whatever the caller asked for (this is a CPU) */
};

static unsigned long long g_stub_calls = 0;

/* homogeneous projective wire point X || Y || Z -> key state like k_prj_load + k_to_affine<MODE 2>:
 * 0 = affine point in P, 1 = point at infinity, -1 = rejected (coordinate >= p or off the curve) */
template <class C> static int import_prj(Aff<C> &P, const uint8_t *b)
{
	typedef Field<typename C::Fp> F;
	constexpr int N = C::N;
	Fe<N> x, y, z, X, Y, Z;
	load_be<N>(x, b, C::PLEN);
	load_be<N>(y, b + C::PLEN, C::PLEN);
	load_be<N>(z, b + 2 * C::PLEN, C::PLEN);
	if (F::geq_mod(x) || F::geq_mod(y) || F::geq_mod(z)) return -1;
	F::to_mont(X, x);
	F::to_mont(Y, y);
	F::to_mont(Z, z);
	if (F::is_zero(Z)) return F::is_zero(X) ? 1 : -1; /* Y^2 * 0 == X^3: X must be 0 (k_prj_load accepts any Y then) */
	Fe<N> zi;
	F::inv(zi, Z);
	F::mul(P.x, X, zi);
	F::mul(P.y, Y, zi);
	return EC<C>::on_curve(P) ? 0 : -1;
}

template <class C> static void store_affine(uint8_t *out, const Aff<C> &P)
{
	typedef Field<typename C::Fp> F;
	Fe<C::N> x, y;
	F::from_mont(x, P.x);
	F::from_mont(y, P.y);
	store_be<C::N>(out, x, C::PLEN);
	store_be<C::N>(out + C::PLEN, y, C::PLEN);
}

extern "C" {

unsigned long long eccb200_stub_calls(void) { return g_stub_calls; }

int eccb200_ctx_create(eccb200_ctx **ctx, int curve_id, int device, int comb_window)
{
	(void)device;
	(void)comb_window;
	static std::mutex mu; /* hostsim builds its comb tables lazily and without locks: build here, serialised */
	std::lock_guard<std::mutex> lk(mu);
	const int w = 4;
	if (!ctx || dispatch(curve_id, [&](auto c) {
		    table_for<decltype(c)>(w);
		    return 0;
	    }) != 0)
		return -1;
	*ctx = new eccb200_ctx{ curve_id, w };
	return 0;
}

void eccb200_ctx_destroy(eccb200_ctx *ctx) { delete ctx; }

void *eccb200_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void *eccb200_host_alloc_input(size_t bytes) { return malloc(bytes ? bytes : 1); }
void eccb200_host_free(void *p) { free(p); }
const char *eccb200_last_error(void) { return "engine stub (tests/hostsim/engine_stub.cpp)"; }

int eccb200_prj_pt_mul_batch(eccb200_ctx *ctx, uint32_t n, const uint8_t *scalars, const uint8_t *points, uint8_t *out,
			     int8_t *status)
{
	if (!ctx) return -1;
	g_stub_calls++;
	return hostsim_prj_pt_mul_batch(ctx->curve_id, ctx->w, n, scalars, points, out, status);
}

int eccb200_prj_pt_unique_batch(eccb200_ctx *ctx, uint32_t n, const uint8_t *prj_points, uint8_t *out_aff, int8_t *status)
{
	if (!ctx) return -1;
	g_stub_calls++;
	return dispatch(ctx->curve_id, [&](auto c) {
		typedef decltype(c) C;
		for (uint32_t i = 0; i < n; i++) {
			Aff<C> P;
			uint8_t *o = out_aff + (size_t)i * 2 * C::PLEN;
			memset(o, 0, 2 * C::PLEN);
			const int ks = import_prj<C>(P, prj_points + (size_t)i * 3 * C::PLEN);
			status[i] = (int8_t)ks;
			if (ks == 0) store_affine<C>(o, P);
		}
		return 0;
	});
}

int eccb200_ecdsa_verify_prj_batch(eccb200_ctx *ctx, uint32_t n, const uint8_t *sigs, const uint8_t *prj_pubkeys,
				   const uint8_t *digests, uint32_t hlen, int8_t *verdict)
{
	if (!ctx) return -1;
	g_stub_calls++;
	const int w = ctx->w;
	return dispatch(ctx->curve_id, [&](auto c) {
		typedef decltype(c) C;
		constexpr int N = C::N;
		const std::vector<uint32_t> &tab = table_for<C>(w);
		for (uint32_t i = 0; i < n; i++) {
			Fe<N> r, s, e, u, v;
			Aff<C> Y;
			load_be<N>(r, sigs + (size_t)i * 2 * C::QLEN, C::QLEN);
			load_be<N>(s, sigs + (size_t)i * 2 * C::QLEN + C::QLEN, C::QLEN);
			const int ks = import_prj<C>(Y, prj_pubkeys + (size_t)i * 3 * C::PLEN);
			verdict[i] = -1;
			if (ks < 0 || !ecdsa_rs_in_range<C>(r, s)) continue;
			if (ks == 1) /* the kernel walks a dummy base for a key at infinity (ecdsa_verify_tail) */
				for (int j = 0; j < N; j++) {
					Y.x.w[j] = C::GX_MONT(j);
					Y.y.w[j] = C::GY_MONT(j);
				}
			digest_to_scalar<C>(e, digests + (size_t)i * hlen, hlen);
			ecdsa_uv<C>(u, v, r, s, e);
			verdict[i] = ecdsa_verify_tail<C>(r, u, v, Y, tab.data(), w, ks == 1) == 0 ? 0 : -1;
		}
		return 0;
	});
}

int eccb200_ecfsdsa_verify_batch(eccb200_ctx *ctx, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				 const uint8_t *digests, uint32_t hlen, int8_t *verdict)
{
	if (!ctx) return -1;
	g_stub_calls++;
	return hostsim_ecfsdsa_verify_batch(ctx->curve_id, ctx->w, n, sigs, pubkeys, digests, hlen, verdict);
}

int eccb200_bip0340_verify_batch(eccb200_ctx *ctx, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				 const uint8_t *digests, uint32_t hlen, int8_t *verdict)
{
	if (!ctx) return -1;
	g_stub_calls++;
	return hostsim_bip0340_verify_batch(ctx->curve_id, ctx->w, n, sigs, pubkeys, digests, hlen, verdict);
}

static const uint8_t kStubSeed[32] = { 0x42, 0x32, 0x30, 0x30, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12,
				       13,   14,   15,   16,   17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28 };

int eccb200_ecfsdsa_verify_msm_batch(eccb200_ctx *ctx, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				     const uint8_t *digests, uint32_t hlen, const uint8_t *seed, int *all_valid)
{
	if (!ctx || !all_valid) return -1;
	g_stub_calls++;
	return hostsim_schnorr_msm(1, ctx->curve_id, 6, n, sigs, pubkeys, digests, hlen, seed ? seed : kStubSeed, all_valid,
				   nullptr);
}

int eccb200_bip0340_verify_msm_batch(eccb200_ctx *ctx, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				     const uint8_t *digests, uint32_t hlen, const uint8_t *seed, int *all_valid)
{
	if (!ctx || !all_valid) return -1;
	g_stub_calls++;
	return hostsim_schnorr_msm(2, ctx->curve_id, 6, n, sigs, pubkeys, digests, hlen, seed ? seed : kStubSeed, all_valid,
				   nullptr);
}

int eccb200_double_smul_batch(eccb200_ctx *ctx, uint32_t n, const uint8_t *ab, const uint8_t *pubkeys, uint8_t *out,
			      int8_t *status)
{
	if (!ctx) return -1;
	g_stub_calls++;
	return hostsim_double_smul_batch(ctx->curve_id, ctx->w, n, ab, pubkeys, out, status);
}

} /* extern "C" */
