/*
 * msm.cuh — K6: Schnorr-type batch verification as one multi-scalar multiplication on the device (the
 * reference: _ecfsdsa_verify_batch sig/ecfsdsa.c:814-1055 with ec_verify_bos_coster sig/sig_algs.c:1052).
 *
 * Stages (all on one stream, nothing returns to the host in between):
 *   prepare     one thread per signature: parse / validate like the reference's loop body, derive the coefficient a_i,
 *               write the points -W_i and Y_i (affine, Montgomery words) with their scalars a_i and a_i e_i mod q, and
 *               reduce a_i s_i mod q over the CTA                                       (k_msm_prepare, k_msm_ssum)
 *   sort        counting sort of (point, window) pairs by bucket: histogram with atomics, tiled exclusive scan, scatter
 *               with atomics; then the buckets themselves are ordered by decreasing load
 *                                    (k_msm_hist, k_msm_scan_*, k_msm_scatter, k_msm_order_hist / k_msm_order_scatter)
 *   accumulate  one thread per bucket: a chain of XYZZ mixed additions (madd-2008-s, 8M + 2S — the addition of the
 *               fixed-base comb) over the bucket's points; this is where the time goes        (k_msm_accumulate)
 *   reduce      running sums over ranges of 16 buckets, a tree per window, Horner over the windows on one thread
 *                                                             (k_msm_reduce, k_msm_window_sum, k_msm_final)
 *
 * HBM layout: pts [2n+1][2N] words, scal [2n+1][N] words, list [<= n (nwin_a + nwin) + nwin] u32 (bit 31 = negate),
 * count / offs / fill [nwin 2^(c-1)] u32, buckets [nwin 2^(c-1)][3N] words, parts [nwin 2^(c-1) / 16][3N], winsum [nwin][3N].
 * Integer work on the multiplier pipe (one accumulate thread issues the same instruction stream as K1's loop body);
 * DRAM traffic is ~26 random 64-byte point reads per signature, far from the HBM roof.
 */
#pragma once
#include "kernels.cuh"
#include "msm_core.cuh"

namespace eccb200 {

struct MsmBuffers {
	uint32_t *pts, *scal, *partial, *count, *offs, *fill, *list, *buckets, *parts, *winsum;
	uint32_t *order; /* [nwin 2^(c-1)] bucket indices by decreasing load */
	uint32_t *aux;   /* [0, 1024) tile totals of the scan, [1024, 2048) load bins, [2048, 3072) their fill counters */
	int *flags; /* [0] = a malformed item was seen, [1] = the verdict (1 = the batch verifies) */
};

/*
 * SCHEME 1, ECFSDSA: sigs [n][2 PLEN + QLEN] = W_x || W_y || s, digests [n][hlen] = H(W_x || W_y || m)
 * SCHEME 2, BIP0340: sigs [n][PLEN + QLEN] = r || s with r = x(R), digests = the tagged challenge hash; R is lifted from r
 *                    and the key to its even-y representative (sig/bip0340.c:1166-1215)
 * pubkeys [n][2 PLEN] affine.
 */
template <class C, int SCHEME>
__global__ void __launch_bounds__(128) k_msm_prepare(uint32_t n, const uint8_t *__restrict__ sigs,
						     const uint8_t *__restrict__ pubkeys,
						     const uint8_t *__restrict__ digests, uint32_t hlen, MsmKey key,
						     int c, uint32_t *__restrict__ pts, uint32_t *__restrict__ scal,
						     uint32_t *__restrict__ partial, int *__restrict__ flags)
{
	typedef Field<typename C::Fp> F;
	typedef Field<typename C::Fq> Fq;
	constexpr int N = C::N;
	__shared__ __align__(16) uint32_t sh[128 * N];
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	Fe<N> t;
	Fq::set_zero(t);
	if (idx < n) {
		Aff<C> W, Y, negW, Yf;
		Fe<N> s, h, a, cY;
		bool w_ok, s_ok, key_ok;
		const uint8_t *pkb = pubkeys + (size_t)idx * (2 * C::PLEN);
		if (SCHEME == 2) {
			const uint8_t *sg = sigs + (size_t)idx * (C::PLEN + C::QLEN);
			Fe<N> r, yraw;
			load_wire<N, C::PLEN>(r, sg);
			load_wire<N, C::QLEN>(s, sg + C::PLEN);
			w_ok = !F::geq_mod(r);                                   /* fp_import_from_buf (sig/bip0340.c:1168) */
			if (!w_ok) F::set_zero(r);
			w_ok = msm_lift_x<C>(W, r) && w_ok;                      /* (:1188-1196) */
			load_wire<N, C::PLEN>(yraw, pkb + C::PLEN);
			key_ok = load_affine_checked<C>(Y, pkb);
			if (yraw.w[0] & 1u) F::neg(Y.y, Y.y);                    /* (:1208-1213) */
		} else {
			const uint8_t *sg = sigs + (size_t)idx * (2 * C::PLEN + C::QLEN);
			w_ok = load_affine_checked<C>(W, sg);                    /* (sig/ecfsdsa.c:983) */
			load_wire<N, C::QLEN>(s, sg + 2 * C::PLEN);
			key_ok = load_affine_checked<C>(Y, pkb);                 /* (:941-942) */
		}
		s_ok = !Fq::geq_mod(s);                                          /* s < q (sig/ecfsdsa.c:919-921, sig/bip0340.c:1171-1172) */
		digest_full_mod_q<C>(h, digests + (size_t)idx * hlen, hlen);     /* (sig/ecfsdsa.c:953-961) */
		Fq::neg(h, h);                                                   /* (:962) */
		msm_coefficient<N>(a, key, idx, c);
		msm_terms<C>(negW, Yf, cY, t, W, Y, s, h, a);
		if (!(w_ok && s_ok && key_ok)) {
			/* the reference returns -1 for the whole batch; the item still owns its slots: zero scalars */
			atomicOr(flags, 1);
			Fq::set_zero(a);
			Fq::set_zero(cY);
			Fq::set_zero(t);
		}
		msm_st<N>(pts + (size_t)idx * (2 * N), negW.x);
		msm_st<N>(pts + (size_t)idx * (2 * N) + N, negW.y);
		msm_st<N>(scal + (size_t)idx * N, a);
		msm_st<N>(pts + ((size_t)n + idx) * (2 * N), Yf.x);
		msm_st<N>(pts + ((size_t)n + idx) * (2 * N) + N, Yf.y);
		msm_st<N>(scal + ((size_t)n + idx) * N, cY);
	}
	/* sum of a_i s_i mod q over the CTA */
#pragma unroll
	for (int j = 0; j < N; j++) sh[threadIdx.x * N + j] = t.w[j];
	for (int step = 64; step > 0; step >>= 1) {
		__syncthreads();
		if ((int)threadIdx.x < step) {
			Fe<N> o;
#pragma unroll
			for (int j = 0; j < N; j++) o.w[j] = sh[(threadIdx.x + step) * N + j];
			Fq::add(t, t, o);
#pragma unroll
			for (int j = 0; j < N; j++) sh[threadIdx.x * N + j] = t.w[j];
		}
	}
	if (threadIdx.x == 0) {
#pragma unroll
		for (int j = 0; j < N; j++) partial[(size_t)blockIdx.x * N + j] = t.w[j];
	}
}

/* the generator's term: point 2n = G with the scalar sum a_i s_i mod q (one CTA) */
template <class C>
__global__ void __launch_bounds__(128) k_msm_ssum(uint32_t nparts, const uint32_t *__restrict__ partial, uint32_t n,
						  uint32_t *__restrict__ pts, uint32_t *__restrict__ scal)
{
	typedef Field<typename C::Fq> Fq;
	constexpr int N = C::N;
	__shared__ __align__(16) uint32_t sh[128 * N];
	Fe<N> t, o;
	Fq::set_zero(t);
	for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x) {
#pragma unroll
		for (int j = 0; j < N; j++) o.w[j] = partial[(size_t)i * N + j];
		Fq::add(t, t, o);
	}
#pragma unroll
	for (int j = 0; j < N; j++) sh[threadIdx.x * N + j] = t.w[j];
	for (int step = 64; step > 0; step >>= 1) {
		__syncthreads();
		if ((int)threadIdx.x < step) {
#pragma unroll
			for (int j = 0; j < N; j++) o.w[j] = sh[(threadIdx.x + step) * N + j];
			Fq::add(t, t, o);
#pragma unroll
			for (int j = 0; j < N; j++) sh[threadIdx.x * N + j] = t.w[j];
		}
	}
	if (threadIdx.x == 0) {
		Fe<N> gx, gy;
#pragma unroll
		for (int j = 0; j < N; j++) {
			gx.w[j] = C::GX_MONT(j);
			gy.w[j] = C::GY_MONT(j);
		}
		if (msm_fold<C>(t)) Field<typename C::Fp>::neg(gy, gy);
		msm_st<N>(pts + (size_t)2 * n * (2 * N), gx);
		msm_st<N>(pts + (size_t)2 * n * (2 * N) + N, gy);
		msm_st<N>(scal + (size_t)2 * n * N, t);
	}
}

template <class C>
__global__ void __launch_bounds__(256) k_msm_hist(uint32_t npts, const uint32_t *__restrict__ scal, int c, int nwin,
						  uint32_t *__restrict__ count)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= npts) return;
	const uint32_t nb = 1u << (c - 1);
	msm_digits(scal + (size_t)j * C::N, C::N, c, nwin, [&](int w, int d) {
		atomicAdd(count + (size_t)w * nb + (uint32_t)((d < 0 ? -d : d) - 1), 1u);
	});
}

/*
 * Exclusive prefix sum of count[0 .. total), total <= 2^20, in three small launches with coalesced accesses: every CTA
 * scans a tile of 1024 entries (four per thread) and publishes its total, one CTA scans the <= 1024 tile totals, and the
 * last kernel adds the tile offsets.
 */
template <class C>
__global__ void __launch_bounds__(256) k_msm_scan_tiles(uint32_t total, const uint32_t *__restrict__ count,
							 uint32_t *__restrict__ offs, uint32_t *__restrict__ tile_sum)
{
	__shared__ uint32_t warp_sum[8];
	const uint32_t base = blockIdx.x * 1024u + threadIdx.x * 4u;
	uint32_t v[4], sum = 0;
#pragma unroll
	for (int i = 0; i < 4; i++) {
		v[i] = base + i < total ? count[base + i] : 0u;
		sum += v[i];
	}
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t incl = sum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += o;
	}
	if (lane == 31) warp_sum[wid] = incl;
	__syncthreads();
	uint32_t before = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) before += w < wid ? warp_sum[w] : 0u;
	uint32_t run = before + incl - sum;
#pragma unroll
	for (int i = 0; i < 4; i++) {
		if (base + i < total) offs[base + i] = run;
		run += v[i];
	}
	if (threadIdx.x == 255) tile_sum[blockIdx.x] = before + incl;
}

/* in-place exclusive scan of n <= 1024 values by one CTA of 1024 threads */
template <class C> __global__ void __launch_bounds__(1024) k_msm_scan_small(uint32_t n, uint32_t *__restrict__ v)
{
	__shared__ uint32_t sh[1024];
	const uint32_t mine = threadIdx.x < n ? v[threadIdx.x] : 0u;
	sh[threadIdx.x] = mine;
	__syncthreads();
	for (int step = 1; step < 1024; step <<= 1) {
		const uint32_t o = (int)threadIdx.x >= step ? sh[threadIdx.x - step] : 0u;
		__syncthreads();
		sh[threadIdx.x] += o;
		__syncthreads();
	}
	if (threadIdx.x < n) v[threadIdx.x] = sh[threadIdx.x] - mine;
}

template <class C>
__global__ void __launch_bounds__(256) k_msm_scan_add(uint32_t total, uint32_t *__restrict__ offs,
						       const uint32_t *__restrict__ tile_off)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < total) offs[i] += tile_off[i >> 10];
}

/*
 * Processing order of the buckets: by decreasing load (counting sort on min(count, 1023)), so that the 32 buckets of a
 * warp of k_msm_accumulate hold (nearly) the same number of points - bucket loads are Poisson distributed and a warp
 * otherwise waits for its fullest bucket (~25 % of the additions idle at a mean load of 32 .. 64).
 */
template <class C>
__global__ void __launch_bounds__(256) k_msm_order_hist(uint32_t total, const uint32_t *__restrict__ count,
							 uint32_t *__restrict__ bins)
{
	__shared__ uint32_t sh[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g < total) {
		const uint32_t cnt = count[g];
		atomicAdd(&sh[1023u - (cnt < 1023u ? cnt : 1023u)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 1024; i += blockDim.x)
		if (sh[i]) atomicAdd(bins + i, sh[i]);
}

template <class C>
__global__ void __launch_bounds__(256) k_msm_order_scatter(uint32_t total, const uint32_t *__restrict__ count,
							    const uint32_t *__restrict__ bin_off,
							    uint32_t *__restrict__ bin_fill, uint32_t *__restrict__ order)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= total) return;
	const uint32_t cnt = count[g], key = 1023u - (cnt < 1023u ? cnt : 1023u);
	order[bin_off[key] + atomicAdd(bin_fill + key, 1u)] = g;
}

template <class C>
__global__ void __launch_bounds__(256) k_msm_scatter(uint32_t npts, const uint32_t *__restrict__ scal, int c, int nwin,
						     const uint32_t *__restrict__ offs, uint32_t *__restrict__ fill,
						     uint32_t *__restrict__ list)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= npts) return;
	const uint32_t nb = 1u << (c - 1);
	msm_digits(scal + (size_t)j * C::N, C::N, c, nwin, [&](int w, int d) {
		const size_t g = (size_t)w * nb + (uint32_t)((d < 0 ? -d : d) - 1);
		const uint32_t pos = offs[g] + atomicAdd(fill + g, 1u);
		list[pos] = j | (d < 0 ? 0x80000000u : 0u);
	});
}

/* one bucket per thread, fullest buckets first: the sum of its points, written as a Jacobian point (Z = 0: empty /
 * cancelled bucket) */
template <class C>
__global__ void __launch_bounds__(128) k_msm_accumulate(uint32_t nbuckets, const uint32_t *__restrict__ order,
							const uint32_t *__restrict__ offs,
							const uint32_t *__restrict__ count,
							const uint32_t *__restrict__ list,
							const uint32_t *__restrict__ pts, uint32_t *__restrict__ buckets)
{
	typedef Field<typename C::Fp> F;
	typedef EC<C> G;
	constexpr int N = C::N;
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nbuckets) return;
	const uint32_t g = order[t], o = offs[g], cnt = count[g];
	typename G::XZ acc;
	G::xz_set_inf(acc);
	for (uint32_t k = 0; k < cnt; k++) {
		const uint32_t e = list[o + k];
		const uint32_t *pp = pts + (size_t)(e & 0x7fffffffu) * (2 * N);
		Aff<C> P;
		msm_ld<N>(P.x, pp);
		msm_ld<N>(P.y, pp + N);
		if (e >> 31) F::neg(P.y, P.y);
		G::xz_add_mixed(acc, acc, P);
	}
	Jac<C> r;
	G::xz_to_jac(r, acc);
	msm_st_jac<C>(buckets, g, r);
}

template <class C>
__global__ void __launch_bounds__(128) k_msm_reduce(uint32_t nparts, uint32_t per_window, uint32_t nb, uint32_t ch,
						    const uint32_t *__restrict__ buckets, uint32_t *__restrict__ parts)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nparts) return;
	const uint32_t w = t / per_window, lo = (t % per_window) * ch;
	Jac<C> r;
	msm_reduce_range<C>(r, buckets, (size_t)w * nb, lo, ch);
	msm_st_jac<C>(parts, t, r);
}

/* one CTA per window: the sum of its per_window partial results */
template <class C>
__global__ void __launch_bounds__(128) k_msm_window_sum(uint32_t per_window, const uint32_t *__restrict__ parts,
							uint32_t *__restrict__ winsum)
{
	typedef EC<C> G;
	constexpr int N = C::N;
	__shared__ __align__(16) uint32_t sh[128 * 3 * N];
	Jac<C> acc, o;
	G::set_inf(acc);
	for (uint32_t i = threadIdx.x; i < per_window; i += blockDim.x) {
		msm_ld_jac<C>(o, parts, (size_t)blockIdx.x * per_window + i);
		G::add_full_ool(acc, acc, o);
	}
	msm_st_jac<C>(sh, threadIdx.x, acc);
	for (int step = 64; step > 0; step >>= 1) {
		__syncthreads();
		if ((int)threadIdx.x < step) {
			msm_ld_jac<C>(o, sh, threadIdx.x + step);
			G::add_full_ool(acc, acc, o);
			msm_st_jac<C>(sh, threadIdx.x, acc);
		}
	}
	if (threadIdx.x == 0) msm_st_jac<C>(winsum, blockIdx.x, acc);
}

/* Horner over the windows, then the verdict: the sum is the point at infinity and no item was malformed */
template <class C>
__global__ void k_msm_final(int nwin, int c, const uint32_t *__restrict__ winsum, int *__restrict__ flags)
{
	if (blockIdx.x || threadIdx.x) return;
	Jac<C> acc;
	msm_horner<C>(acc, winsum, nwin, c);
	flags[1] = (EC<C>::is_inf(acc) && flags[0] == 0) ? 1 : 0;
}

template <class C> struct LaunchMsm {
	/* enqueues the whole verification of n signatures (scheme 1 ECFSDSA, 2 BIP0340); flags[1] holds the verdict when the
	 * stream drains.  Returns the number of kernels launched, -1 when the scheme cannot run on this curve (BIP0340 needs
	 * p = 3 mod 4 for the lift of r). */
	static int verify(int scheme, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys, const uint8_t *digests,
			  uint32_t hlen, const MsmKey &key, int c, const MsmBuffers &b, cudaStream_t st);
};

#if defined(ECC_TU_MSM)
template <class C>
int LaunchMsm<C>::verify(int scheme, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys, const uint8_t *digests,
			 uint32_t hlen, const MsmKey &key, int c, const MsmBuffers &b, cudaStream_t st)
{
	if (scheme == 2 && !msm_lift_supported<C>()) return -1;
	const int nwin = msm_windows(C::QBITS - 1, c);
	const uint32_t nb = 1u << (c - 1), total = (uint32_t)nwin * nb, ch = nb < 16u ? nb : 16u, per_window = nb / ch,
		       nparts = (uint32_t)nwin * per_window, npts = 2 * n + 1, nblk = (n + 127) / 128,
		       tiles = (total + 1023u) / 1024u;
	cudaMemsetAsync(b.count, 0, (size_t)total * 4, st);
	cudaMemsetAsync(b.fill, 0, (size_t)total * 4, st);
	cudaMemsetAsync(b.aux, 0, 3072 * 4, st);
	cudaMemsetAsync(b.flags, 0, 2 * sizeof(int), st);
	if (scheme == 2) {
		if constexpr (msm_lift_supported<C>())
			k_msm_prepare<C, 2><<<nblk, 128, 0, st>>>(n, sigs, pubkeys, digests, hlen, key, c, b.pts, b.scal, b.partial,
								   b.flags);
	} else {
		k_msm_prepare<C, 1><<<nblk, 128, 0, st>>>(n, sigs, pubkeys, digests, hlen, key, c, b.pts, b.scal, b.partial, b.flags);
	}
	k_msm_ssum<C><<<1, 128, 0, st>>>(nblk, b.partial, n, b.pts, b.scal);
	k_msm_hist<C><<<(npts + 255) / 256, 256, 0, st>>>(npts, b.scal, c, nwin, b.count);
	k_msm_scan_tiles<C><<<tiles, 256, 0, st>>>(total, b.count, b.offs, b.aux);
	k_msm_scan_small<C><<<1, 1024, 0, st>>>(tiles, b.aux);
	k_msm_scan_add<C><<<(total + 255) / 256, 256, 0, st>>>(total, b.offs, b.aux);
	k_msm_order_hist<C><<<(total + 255) / 256, 256, 0, st>>>(total, b.count, b.aux + 1024);
	k_msm_scan_small<C><<<1, 1024, 0, st>>>(1024, b.aux + 1024);
	k_msm_order_scatter<C><<<(total + 255) / 256, 256, 0, st>>>(total, b.count, b.aux + 1024, b.aux + 2048, b.order);
	k_msm_scatter<C><<<(npts + 255) / 256, 256, 0, st>>>(npts, b.scal, c, nwin, b.offs, b.fill, b.list);
	k_msm_accumulate<C><<<(total + 127) / 128, 128, 0, st>>>(total, b.order, b.offs, b.count, b.list, b.pts, b.buckets);
	k_msm_reduce<C><<<(nparts + 127) / 128, 128, 0, st>>>(nparts, per_window, nb, ch, b.buckets, b.parts);
	k_msm_window_sum<C><<<nwin, 128, 0, st>>>(per_window, b.parts, b.winsum);
	k_msm_final<C><<<1, 32, 0, st>>>(nwin, c, b.winsum, b.flags);
	return 14;
}
#endif

} // namespace eccb200
