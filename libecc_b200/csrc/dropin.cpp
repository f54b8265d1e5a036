/*
 * dropin.cpp — include/libecc_b200_dropin.h: the reference's own entry points (prj_pt_mul, prj_pt_mul_blind, the
 * ECDSA verify_batch slot) on the reference's own structs, forwarding to the GPU engine (libecc_b200.h).
 *
 * Host work done here is marshalling only: struct validation (magic words), curve identification, the scalar
 * reduction m mod order that the reference's ladder performs implicitly (curves/prj_pt.c:1591-1619; SURVEY.md §8a:
 * "host shim should reduce k mod crv->order"), byte-order conversion, and filling valid output structs.  All field
 * and group arithmetic runs on the device.
 */
#include "../../include/libecc_b200.h"
#include "../../include/libecc_b200_dropin.h"
#include "fp.cuh" /* curve constants only (host build: nothing here is executed as arithmetic) */

#include <dlfcn.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <string>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

using namespace eccb200;

namespace {

constexpr uint64_t kPrjPtMagic = 0xe1cd70babb1d5afeULL;  /* curves/prj_pt.c:26 */
constexpr uint64_t kFpMagic = 0x14e96c8ab28221efULL;     /* fp/fp.c:127 */
constexpr uint64_t kNnMagic = 0xb4cf5d56e2023316ULL ^ (uint64_t)(ECCB200_NN_MAX_WORD_LEN + 64); /* nn/nn.c:28 */
constexpr uint64_t kPubKeyMagic = 0x31327f37741ffb76ULL; /* sig/ec_key.h:118 */
constexpr int kMaxWords = ECCB200_NN_MAX_WORD_LEN;
constexpr int kNumCurves = 11;

struct CurveInfo {
	int id;
	int n64;              /* 64-bit limbs of p / q */
	int plen, qlen;       /* wire bytes of a field element / of a scalar (BYTECEIL of the bit lengths) */
	uint64_t p[9], q[9], gx[9], gy[9];
};

template <class C> CurveInfo make_info()
{
	CurveInfo ci;
	memset(&ci, 0, sizeof(ci));
	ci.id = C::ID;
	ci.n64 = C::N / 2;
	ci.plen = C::PLEN;
	ci.qlen = C::QLEN;
	for (int i = 0; i < C::N / 2; i++) {
		ci.p[i] = ((uint64_t)C::Fp::P(2 * i + 1) << 32) | C::Fp::P(2 * i);
		ci.q[i] = ((uint64_t)C::Fq::P(2 * i + 1) << 32) | C::Fq::P(2 * i);
		ci.gx[i] = ((uint64_t)C::GX(2 * i + 1) << 32) | C::GX(2 * i);
		ci.gy[i] = ((uint64_t)C::GY(2 * i + 1) << 32) | C::GY(2 * i);
	}
	return ci;
}

const CurveInfo *curves()
{
	static const CurveInfo tab[kNumCurves] = { make_info<Curve_SECP256R1>(),       make_info<Curve_FRP256V1>(),
						    make_info<Curve_SECP384R1>(),       make_info<Curve_BRAINPOOLP256R1>(),
						    make_info<Curve_BRAINPOOLP384R1>(), make_info<Curve_SECP256K1>(),
						    make_info<Curve_SECP521R1>(),
						    make_info<Curve_SM2P256V1>(),
						    make_info<Curve_BRAINPOOLP512R1>(),
						    make_info<Curve_SECP224R1>(),
						    make_info<Curve_SECP192R1>() };
	return tab;
}

bool nn_ok(const eccb200_nn *a) { return a && a->magic == kNnMagic && a->wlen <= kMaxWords; }
bool fp_ok(const eccb200_fp *a) { return a && a->magic == kFpMagic && a->ctx && nn_ok(&a->fp_val); }
bool pt_ok(const eccb200_prj_pt *p)
{
	return p && p->magic == kPrjPtMagic && p->crv && fp_ok(&p->X) && fp_ok(&p->Y) && fp_ok(&p->Z);
}

bool words_eq(const eccb200_nn *a, const uint64_t *w, int n)
{
	for (int i = 0; i < kMaxWords; i++)
		if (a->val[i] != (i < n ? w[i] : 0)) return false;
	return true;
}

const CurveInfo *identify(const eccb200_prj_pt *pt)
{
	const eccb200_fp_ctx *ctx = pt->X.ctx;
	if (!nn_ok(&ctx->p) || !nn_ok(&pt->crv->order)) return nullptr;
	for (int c = 0; c < kNumCurves; c++) {
		const CurveInfo *ci = &curves()[c];
		if (words_eq(&ctx->p, ci->p, ci->n64) && words_eq(&pt->crv->order, ci->q, ci->n64)) return ci;
	}
	return nullptr;
}

/*
 * Engine contexts.  The reference's functions are re-entrant and lock-free; an eccb200_ctx serves one thread at a
 * time.  So the layer keeps, per curve, a few SMALL contexts (16-bit comb: 64 MiB table for a 256-bit curve, built in
 * tens of milliseconds) that concurrent single calls / small batches pick from without blocking each other, and ONE
 * BIG context (the library's default 22-bit comb, 3.2 GiB for a 256-bit curve; ECCB200_COMB_WINDOW overrides) that is
 * only created when a batch of at least kBigBatch items arrives.  Nothing is created for curves that are never used;
 * eccb200_dropin_release() frees everything.
 */
constexpr uint32_t kMaxBatch = 1u << 28; /* items per call: keeps every 32-bit size product below 2^32 (a batch of that
                                         * size is 210 GB of ec_pub_key structs - split it) */
constexpr int kSmallSlots = 4;
constexpr uint32_t kBigBatch = 1u << 15;

struct Staging { /* page-locked staging of a slot, grown on demand (DMA'd directly by the engine's pipeline) */
	uint8_t *p = nullptr;
	size_t cap = 0;
	uint8_t *get(size_t bytes)
	{
		if (bytes <= cap) return p;
		if (p) eccb200_host_free(p);
		cap = 0;
		p = (uint8_t *)eccb200_host_alloc(bytes);
		if (p) cap = bytes;
		return p;
	}
	void release()
	{
		if (p) eccb200_host_free(p);
		p = nullptr;
		cap = 0;
	}
};

struct Slot {
	std::mutex mu;
	eccb200_ctx *ctx = nullptr;
	Staging st[6];
};

struct CurveEngines {
	Slot small_[kSmallSlots];
	Slot big;
};

CurveEngines g_eng[32]; /* indexed by the reference's ec_curve_type (< 32 here) */
std::atomic<unsigned long long> g_calls{ 0 };   /* scalar multiplications served (eccb200_dropin_call_count) */
std::atomic<unsigned long long> g_verifies{ 0 }; /* signatures verified on the GPU (eccb200_dropin_verify_count) */
std::atomic<unsigned long long> g_msm_batches{ 0 }; /* batches settled by the multi-scalar-multiplication fast path */
std::atomic<int> g_device{ -1 };
std::atomic<int> g_blind_on_gpu{ -1 };

int device_index()
{
	int d = g_device.load();
	if (d < 0) {
		const char *e = getenv("ECCB200_DEVICE");
		d = e ? atoi(e) : 0;
		g_device.store(d);
	}
	return d;
}

/* RAII: a locked slot with a live context (ctx == nullptr when creation failed) */
struct Engine {
	Slot *slot = nullptr;
	eccb200_ctx *ctx = nullptr;
	std::unique_lock<std::mutex> lk;
};

Engine acquire(int curve_id, uint64_t n)
{
	Engine e;
	CurveEngines &ce = g_eng[curve_id];
	int w = 16;
	if (n >= kBigBatch) {
		e.slot = &ce.big;
		e.lk = std::unique_lock<std::mutex>(ce.big.mu);
		const char *ws = getenv("ECCB200_COMB_WINDOW");
		w = ws ? atoi(ws) : 0;
	} else {
		for (int i = 0; i < kSmallSlots && !e.slot; i++) {
			std::unique_lock<std::mutex> l(ce.small_[i].mu, std::try_to_lock);
			if (l.owns_lock()) {
				e.slot = &ce.small_[i];
				e.lk = std::move(l);
			}
		}
		if (!e.slot) { /* all busy: queue on the slot this thread hashes to */
			size_t h = std::hash<std::thread::id>()(std::this_thread::get_id());
			e.slot = &ce.small_[h % kSmallSlots];
			e.lk = std::unique_lock<std::mutex>(e.slot->mu);
		}
		const char *ws = getenv("ECCB200_DROPIN_SMALL_WINDOW");
		if (ws && atoi(ws) >= 4 && atoi(ws) <= 16) w = atoi(ws);
	}
	if (!e.slot->ctx && eccb200_ctx_create(&e.slot->ctx, curve_id, device_index(), w)) e.slot->ctx = nullptr;
	e.ctx = e.slot->ctx;
	return e;
}

void release_all()
{
	for (auto &ce : g_eng) {
		Slot *all[kSmallSlots + 1];
		for (int i = 0; i < kSmallSlots; i++) all[i] = &ce.small_[i];
		all[kSmallSlots] = &ce.big;
		for (Slot *s : all) {
			std::lock_guard<std::mutex> lk(s->mu);
			if (s->ctx) eccb200_ctx_destroy(s->ctx);
			s->ctx = nullptr;
			for (auto &b : s->st) b.release();
		}
	}
}

/* host threads for the marshalling loops of the batch entry points */
template <class F> void parallel_for(uint32_t n, F f)
{
	unsigned hw = std::thread::hardware_concurrency(), cap = 32u;
	if (const char *e = getenv("ECCB200_DROPIN_THREADS")) { /* an explicit request may exceed the default cap of 32 */
		hw = (unsigned)atoi(e);
		cap = 128u;
	}
	unsigned T = std::min<unsigned>(std::min<unsigned>(hw ? hw : 1, cap), n / 2048u + 1u);
	if (T <= 1) {
		f(0u, n, 0u);
		return;
	}
	std::vector<std::thread> th;
	for (unsigned t = 0; t < T; t++) {
		uint32_t lo = (uint32_t)((uint64_t)n * t / T), hi = (uint32_t)((uint64_t)n * (t + 1) / T);
		th.emplace_back([=] { f(lo, hi, t); });
	}
	for (auto &x : th) x.join();
}

/* little-endian 64-bit words -> len big-endian bytes (nn_export_to_buf, nn/nn.c:511) */
void words_to_be(uint8_t *out, int len, const uint64_t *w)
{
	for (int j = 0; j < len; j++) out[len - 1 - j] = (uint8_t)(w[j / 8] >> (8 * (j % 8)));
}

/* m mod q -> big-endian qlen bytes.  Plain binary long division on 64-bit words (nn_mod, nn/nn_div.c:1005). */
void scalar_mod_to_be(uint8_t *out, const eccb200_nn *m, const CurveInfo *ci)
{
	const int n = ci->n64;
	uint64_t r[10] = { 0 };
	int top = m->wlen > kMaxWords ? kMaxWords : m->wlen;
	for (int wi = top - 1; wi >= 0; wi--) {
		for (int b = 63; b >= 0; b--) {
			/* r = 2r + bit */
			uint64_t carry = (m->val[wi] >> b) & 1;
			for (int i = 0; i <= n; i++) {
				uint64_t nc = r[i] >> 63;
				r[i] = (r[i] << 1) | carry;
				carry = nc;
			}
			/* if r >= q: r -= q */
			bool ge = r[n] != 0;
			if (!ge) {
				ge = true;
				for (int i = n - 1; i >= 0; i--) {
					if (r[i] != ci->q[i]) {
						ge = r[i] > ci->q[i];
						break;
					}
				}
			}
			if (ge) {
				unsigned __int128 bw = 0;
				for (int i = 0; i <= n; i++) {
					unsigned __int128 t = (unsigned __int128)r[i] - (i < n ? ci->q[i] : 0) - (uint64_t)bw;
					r[i] = (uint64_t)t;
					bw = (t >> 64) & 1;
				}
			}
		}
	}
	words_to_be(out, ci->qlen, r);
}

void fp_to_be(uint8_t *out, const eccb200_fp *a, int len) { words_to_be(out, len, a->fp_val.val); }

void gen_to_be(uint8_t *pp, const CurveInfo *ci)
{
	words_to_be(pp, ci->plen, ci->gx);
	words_to_be(pp + ci->plen, ci->plen, ci->gy);
}

bool fp_is_small(const eccb200_fp *a, uint64_t v)
{
	if (a->fp_val.val[0] != v) return false;
	for (int i = 1; i < kMaxWords; i++)
		if (a->fp_val.val[i]) return false;
	return true;
}

void fp_set_be(eccb200_fp *dst, const eccb200_fp *tmpl, const uint8_t *be, int len)
{
	*dst = *tmpl; /* ctx pointer, magics */
	memset(dst->fp_val.val, 0, sizeof(dst->fp_val.val));
	if (be) {
		for (int j = 0; j < len; j++) dst->fp_val.val[j / 8] |= (uint64_t)be[len - 1 - j] << (8 * (j % 8));
	}
	dst->fp_val.magic = kNnMagic;
	dst->fp_val.wlen = tmpl->ctx->p.wlen; /* elements carry the word length of p (fp_init, fp/fp.c:139) */
	dst->magic = kFpMagic;
}

void fp_set_word(eccb200_fp *dst, const eccb200_fp *tmpl, uint64_t v)
{
	fp_set_be(dst, tmpl, nullptr, 0);
	dst->fp_val.val[0] = v;
}

/* core: out[i] = m[i]*in[i]; returns 0 iff all ok */
int mul_batch(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in, uint32_t n, int *ret)
{
	if (n == 0) return 0;
	if (!out || !m || !in || n > kMaxBatch) return -1;
	g_calls += n;
	std::vector<int> rc(n, -1);
	const CurveInfo *ci = nullptr;
	for (uint32_t i = 0; i < n; i++) {
		if (!pt_ok(&in[i]) || !nn_ok(&m[i])) continue;
		const CurveInfo *c = identify(&in[i]);
		if (!c || (ci && c != ci)) continue;
		ci = c;
		rc[i] = 0;
	}
	int all = 0;
	if (!ci) {
		if (ret) memcpy(ret, rc.data(), n * sizeof(int));
		return -1;
	}
	auto engine_failed = [&]() { /* no per-item result exists: every item reports -1 */
		if (ret)
			for (uint32_t i = 0; i < n; i++) ret[i] = -1;
		return -1;
	};
	Engine engine = acquire(ci->id, n);
	eccb200_ctx *eng = engine.ctx;
	if (!eng) return engine_failed();
	const size_t plen = (size_t)ci->plen, qlen = (size_t)ci->qlen;
	const int pl = ci->plen;
	std::vector<uint8_t> scalars(n * qlen), points(n * 2 * plen), outb(n * 2 * plen);
	std::vector<int8_t> status(n);
	/* 1. inputs with Z != 1 are normalised on the device first (batched prj_pt_unique) */
	std::vector<uint32_t> prj_idx;
	bool all_gen = true;
	for (uint32_t i = 0; i < n; i++) {
		if (rc[i]) continue;
		if (!fp_is_small(&in[i].Z, 1)) prj_idx.push_back(i);
	}
	std::vector<uint8_t> is_inf(n, 0);
	if (!prj_idx.empty()) {
		std::vector<uint8_t> pb(prj_idx.size() * 3 * plen), ab(prj_idx.size() * 2 * plen);
		std::vector<int8_t> st(prj_idx.size());
		for (size_t k = 0; k < prj_idx.size(); k++) {
			const eccb200_prj_pt *p = &in[prj_idx[k]];
			fp_to_be(&pb[k * 3 * plen], &p->X, pl);
			fp_to_be(&pb[k * 3 * plen + plen], &p->Y, pl);
			fp_to_be(&pb[k * 3 * plen + 2 * plen], &p->Z, pl);
		}
		if (eccb200_prj_pt_unique_batch(eng, (uint32_t)prj_idx.size(), pb.data(), ab.data(), st.data())) return engine_failed();
		for (size_t k = 0; k < prj_idx.size(); k++) {
			uint32_t i = prj_idx[k];
			if (st[k] < 0) rc[i] = -1;                      /* not on the curve: prj_pt_mul fails (:1767) */
			else if (st[k] == 1) is_inf[i] = 1;             /* in = infinity -> out = infinity, ret 0 */
			else memcpy(&points[i * 2 * plen], &ab[k * 2 * plen], 2 * plen);
		}
	}
	for (uint32_t i = 0; i < n; i++) {
		if (rc[i]) {
			memset(&scalars[i * qlen], 0, qlen);
			continue;
		}
		scalar_mod_to_be(&scalars[i * qlen], &m[i], ci);
		if (fp_is_small(&in[i].Z, 1)) {
			fp_to_be(&points[i * 2 * plen], &in[i].X, pl);
			fp_to_be(&points[i * 2 * plen + plen], &in[i].Y, pl);
		}
		if (is_inf[i]) { /* any valid point works as a placeholder; the result is forced to infinity below */
			gen_to_be(&points[i * 2 * plen], ci);
			memset(&scalars[i * qlen], 0, qlen);
		}
		/* fixed-base fast path only when every base is the generator */
		if (all_gen) {
			uint8_t gb[2 * 72];
			gen_to_be(gb, ci);
			if (memcmp(gb, &points[i * 2 * plen], 2 * plen)) all_gen = false;
		}
	}
	for (uint32_t i = 0; i < n; i++)
		if (rc[i]) gen_to_be(&points[i * 2 * plen], ci); /* keep the batch launchable: rejected slots get G */
	if (eccb200_prj_pt_mul_batch(eng, n, scalars.data(), all_gen ? nullptr : points.data(), outb.data(),
				     status.data()))
		return engine_failed();
	for (uint32_t i = 0; i < n; i++) {
		if (rc[i] == 0 && status[i] < 0) rc[i] = -1;
		if (rc[i]) {
			all = -1;
			continue;
		}
		const eccb200_prj_pt src = in[i]; /* copy first: out may alias in (curves/prj_pt.c:1769) */
		eccb200_prj_pt *o = &out[i];
		if (status[i] == 1 || is_inf[i]) { /* canonical infinity (0, 1, 0), prj_pt_zero curves/prj_pt.c:124-136 */
			fp_set_word(&o->X, &src.X, 0);
			fp_set_word(&o->Y, &src.Y, 1);
			fp_set_word(&o->Z, &src.Z, 0);
		} else {
			fp_set_be(&o->X, &src.X, &outb[i * 2 * plen], pl);
			fp_set_be(&o->Y, &src.Y, &outb[i * 2 * plen + plen], pl);
			fp_set_word(&o->Z, &src.Z, 1);
		}
		o->crv = src.crv;
		o->magic = kPrjPtMagic;
	}
	if (ret) memcpy(ret, rc.data(), n * sizeof(int));
	return all;
}

/* head of the reference's hash_mapping (hash/hash_algs.h:232-241) */
struct HashMappingHead {
	int type;
	const char *name;
	uint8_t digest_size;
	uint8_t block_size;
	void *hfunc_init;
	void *hfunc_update;
	void *hfunc_finalize;
	int (*hfunc_scattered)(const unsigned char **inputs, const uint32_t *ilens, unsigned char *output);
};

thread_local std::vector<int8_t> t_verdicts;

} // namespace

extern "C" int eccb200_dropin_set_device(int device)
{
	release_all();
	g_device.store(device);
	return 0;
}

extern "C" void eccb200_dropin_release(void) { release_all(); }

extern "C" int eccb200_dropin_prj_pt_mul_batch(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in,
					       uint32_t n, int *ret)
{
	return mul_batch(out, m, in, n, ret);
}

extern "C" int eccb200_dropin_prj_pt_mul(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in)
{
	return mul_batch(out, m, in, 1, nullptr);
}

extern "C" int prj_pt_mul(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in)
{
	return mul_batch(out, m, in, 1, nullptr);
}

/*
 * prj_pt_mul_blind (curves/prj_pt.c:1782-1822) exists to protect a SECRET scalar (the ECDSA nonce under
 * USE_SIG_BLINDING, sig/ecdsa_common.c:476): it blinds the scalar and runs the constant-time ladder.  The GPU path is
 * a throughput path — its table indices and branches depend on the scalar — so the drop-in does NOT silently take such
 * calls over: by default the call is forwarded to the next definition of prj_pt_mul_blind in the process (the
 * reference's own, when the drop-in is preloaded or linked ahead of a shared libec), and fails with -1 if there is
 * none.  ECCB200_BLIND_ON_GPU=1 (or eccb200_dropin_allow_nonct_blind(1)) opts in to the GPU path, which returns the
 * same point without the side-channel protection.
 */
typedef int (*mul_sig)(eccb200_prj_pt *, const eccb200_nn *, const eccb200_prj_pt *);

/* Is `fn` defined by THIS shared object?  (Comparing with &our_function is not enough: a reference to an exported
 * function of a -fPIC library is itself bound through the global lookup scope, i.e. possibly to the reference's.) */
static void self_anchor() {}
static bool defined_here(void *fn)
{
	Dl_info a, b;
	if (!fn || !dladdr(fn, &a) || !dladdr((void *)&self_anchor, &b)) return false;
	return a.dli_fbase == b.dli_fbase;
}

/* the next definition of `name` after this library: RTLD_NEXT when preloaded / linked ahead, else (library opened
 * privately with dlopen) whatever the global scope holds, as long as it is not our own */
static void *next_definition(const char *name)
{
	void *f = dlsym(RTLD_NEXT, name);
	if (f && !defined_here(f)) return f;
	f = dlsym(RTLD_DEFAULT, name);
	return (f && !defined_here(f)) ? f : nullptr;
}

extern "C" int prj_pt_mul_blind(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in);

static bool blind_on_gpu()
{
	int v = g_blind_on_gpu.load();
	if (v < 0) {
		const char *e = getenv("ECCB200_BLIND_ON_GPU");
		v = (e && atoi(e) != 0) ? 1 : 0;
		g_blind_on_gpu.store(v);
	}
	return v == 1;
}

extern "C" void eccb200_dropin_allow_nonct_blind(int on) { g_blind_on_gpu.store(on ? 1 : 0); }

extern "C" int prj_pt_mul_blind(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in)
{
	if (blind_on_gpu()) return mul_batch(out, m, in, 1, nullptr);
	static mul_sig next = (mul_sig)next_definition("prj_pt_mul_blind");
	return next ? next(out, m, in) : -1;
}

extern "C" unsigned long long eccb200_dropin_call_count(void) { return g_calls.load(); }
extern "C" unsigned long long eccb200_dropin_msm_batches(void) { return g_msm_batches.load(); }

extern "C" unsigned long long eccb200_dropin_verify_count(void) { return g_verifies.load(); }

extern "C" uint32_t eccb200_dropin_last_verdicts(int8_t *verdicts, uint32_t cap)
{
	uint32_t k = (uint32_t)t_verdicts.size();
	if (k > cap) k = cap;
	if (verdicts && k) memcpy(verdicts, t_verdicts.data(), k);
	return k;
}

typedef int (*get_hash_fn)(int, const HashMappingHead **);
static get_hash_fn resolve_get_hash()
{
	/* the reference's own hash (src/hash stays host-side): get_hash_by_type from the application / libsign */
	static get_hash_fn f = (get_hash_fn)dlsym(RTLD_DEFAULT, "get_hash_by_type");
	return f;
}

/* The schemes served by the verification kernel and their ec_alg_type values (lib_ecc_types.h:22-80). */
enum Scheme { kEcdsa = 0, kEcfsdsa = 1, kBip0340 = 2 };
static bool scheme_of(int sig_type, Scheme *sc)
{
	if (sig_type == 1 || sig_type == 14) *sc = kEcdsa; /* ECDSA, DECDSA */
	else if (sig_type == 5) *sc = kEcfsdsa;
	else if (sig_type == 20) *sc = kBip0340;
	else return false;
	return true;
}

/* batches settled by the multi-scalar-multiplication fast path (eccb200_dropin_msm_batches) and its threshold: below
 * it the fixed cost of the bucket method (14 launches, a serial Horner tail of ~1 ms) outweighs one wave of the per-item
 * kernel; ECCB200_DROPIN_MSM_MIN overrides (0 disables the fast path). */
static uint32_t msm_min_batch()
{
	static const uint32_t v = [] {
		const char *e = getenv("ECCB200_DROPIN_MSM_MIN");
		if (!e) return 16384u;
		const long long x = atoll(e);
		return x <= 0 ? 0xffffffffu : (uint32_t)std::min<long long>(x, 0xffffffffll);
	}();
	return v;
}

/*
 * One batch through the verification kernel.
 *   ECDSA / DECDSA  signature r || s,               digest H(m)
 *   ECFSDSA         signature W_x || W_y || s,      digest H(W_x || W_y || m)            (sig/ecfsdsa.c:482,529)
 *   BIP0340         signature r || s (r = x(kG)),   digest H(H(tag) || H(tag) || r || x(Y) || m), tag = "BIP0340/challenge"
 *                                                                                         (sig/bip0340.c:45-69,438-443)
 * Hashing uses the reference's own src/hash (hfunc_scattered of the hash mapping).
 */
static int verify_batch_common(Scheme sc, const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
			       const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
			       const uint8_t **adata)
{
	if (num > kMaxBatch) {
		t_verdicts.clear();
		return -1;
	}
	t_verdicts.assign(num, -1);
	if (num == 0) return -1; /* the reference's implementations reject an empty batch (sig/ecfsdsa.c:740) */
	if (!s || !s_len || !pub_keys || !m || !m_len) return -1;
	if (adata) /* none of the three schemes takes ancillary data: every entry must be NULL (sig/ecdsa.c:76-83) */
		for (uint32_t i = 0; i < num; i++)
			if (adata[i]) return -1;
	get_hash_fn get_hash = resolve_get_hash();
	if (!get_hash) return -1;
	const HashMappingHead *hm = nullptr;
	if (get_hash(hash_type, &hm) || !hm || !hm->hfunc_scattered) return -1;
	const uint32_t hlen = hm->digest_size;
	if (hlen == 0 || hlen > 128) return -1;

	/* the curve: the first key that identifies one; every other key must agree (sig/ecfsdsa.c:711) */
	const CurveInfo *ci = nullptr;
	for (uint32_t i = 0; i < num && !ci; i++) {
		const eccb200_ec_pub_key *pk = pub_keys[i];
		if (pk && pk->magic == kPubKeyMagic && pk->key_type == sig_type && pt_ok(&pk->y)) ci = identify(&pk->y);
	}
	if (!ci) return -1;
	Engine engine = acquire(ci->id, num);
	eccb200_ctx *eng = engine.ctx;
	if (!eng) return -1;
	const int pl = ci->plen;
	const size_t plen = (size_t)ci->plen, qlen = (size_t)ci->qlen;
	const size_t siglen = sc == kEcfsdsa ? 2 * plen + qlen : (sc == kBip0340 ? plen + qlen : 2 * qlen);
	/* ECDSA: the keys travel in the reference's projective form (X || Y || Z) and are normalised on the device in front
	 * of the verification kernel; the other two take affine keys (Z != 1 ones go through eccb200_prj_pt_unique_batch) */
	const size_t keylen = sc == kEcdsa ? 3 * plen : 2 * plen;
	/* page-locked staging owned by the slot: the engine's pipeline DMAs straight out of / into it */
	uint8_t *sigs = engine.slot->st[0].get(num * siglen), *pubs = engine.slot->st[1].get(num * keylen),
		*dig = engine.slot->st[2].get(num * (size_t)hlen);
	int8_t *verdict = (int8_t *)engine.slot->st[3].get(num);
	if (!sigs || !pubs || !dig || !verdict) return -1;
	std::vector<uint8_t> ok(num, 0);
	std::atomic<int> mixed{ 0 }, any_prj{ 0 };
	/* ECFSDSA / BIP0340 take affine keys; an ec_pub_key made by ec_key_pair_gen holds a projective point with Z != 1.
	 * Those batches write X || Y || Z for every item into a page-locked staging buffer and go through ONE batched
	 * prj_pt_unique on the device (pipelined DMA in and out of the staging) before the verification. */
	uint8_t *prj = nullptr;
	int8_t *prj_st = nullptr;
	if (sc != kEcdsa) {
		prj = engine.slot->st[4].get(num * 3 * plen);
		prj_st = (int8_t *)engine.slot->st[5].get(num);
		if (!prj || !prj_st) return -1;
	}
	/* marshalling: struct checks, byte-order conversion and the reference's own hash, on several host threads */
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
		bool saw_prj = false;
		for (uint32_t i = lo; i < hi; i++) {
			memset(&sigs[i * siglen], 0, siglen);
			memset(&pubs[i * keylen], 0, keylen);
			memset(&dig[i * (size_t)hlen], 0, hlen);
			if (prj) memset(&prj[i * 3 * plen], 0, 3 * plen);
			const eccb200_ec_pub_key *pk = pub_keys[i];
			if (!pk || pk->magic != kPubKeyMagic || pk->key_type != sig_type || !pt_ok(&pk->y)) continue;
			if (!s[i] || (!m[i] && m_len[i])) continue;
			const CurveInfo *c = identify(&pk->y);
			if (!c) continue;
			if (c != ci) {
				mixed.store(1);
				continue;
			}
			if (s_len[i] != siglen) continue; /* siglen check, sig/ecdsa_common.c:645, sig/ecfsdsa.c:447, sig/bip0340.c:421 */
			memcpy(&sigs[i * siglen], s[i], siglen);
			if (sc != kBip0340) {
				const bool fs = sc == kEcfsdsa;
				const unsigned char *inputs[3] = { fs ? s[i] : m[i], fs ? m[i] : nullptr, nullptr };
				uint32_t ilens[2] = { fs ? (uint32_t)(2 * plen) : m_len[i], fs ? m_len[i] : 0 };
				if (hm->hfunc_scattered(inputs, ilens, &dig[i * (size_t)hlen])) continue;
			}
			const eccb200_prj_pt *y = &pk->y;
			if (sc == kEcdsa) {
				fp_to_be(&pubs[i * keylen], &y->X, pl);
				fp_to_be(&pubs[i * keylen + plen], &y->Y, pl);
				fp_to_be(&pubs[i * keylen + 2 * plen], &y->Z, pl);
			} else {
				const bool affine = fp_is_small(&y->Z, 1);
				if (affine) {
					fp_to_be(&pubs[i * 2 * plen], &y->X, pl);
					fp_to_be(&pubs[i * 2 * plen + plen], &y->Y, pl);
				}
				saw_prj = saw_prj || !affine;
				fp_to_be(&prj[i * 3 * plen], &y->X, pl);
				fp_to_be(&prj[i * 3 * plen + plen], &y->Y, pl);
				fp_to_be(&prj[i * 3 * plen + 2 * plen], &y->Z, pl);
			}
			ok[i] = 1;
		}
		if (saw_prj) any_prj.store(1);
	});
	if (mixed.load()) return -1; /* all keys must share the curve parameters */
	if (sc != kEcdsa && any_prj.load()) {
		/* every item's key through the batched prj_pt_unique (items refused above carry Z = 0 and are ignored) */
		if (eccb200_prj_pt_unique_batch(eng, num, prj, pubs, prj_st)) return -1;
		parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
			for (uint32_t i = lo; i < hi; i++) {
				if (!ok[i] || prj_st[i] == 0) continue;
				if (sc == kEcfsdsa && prj_st[i] == 1) {
					/* ECFSDSA, key at infinity: the reference goes on with e*Y = infinity, W' = s*G
					 * (sig/ecfsdsa.c:594-600 on the complete formulas) and accepts iff s*G == r.  Same here: the
					 * item runs with e = 0 (an all-zero digest) on a dummy base. */
					gen_to_be(&pubs[i * 2 * plen], ci);
					memset(&dig[i * (size_t)hlen], 0, hlen);
					continue;
				}
				ok[i] = 0; /* off the curve; or BIP0340 with a key at infinity: prj_pt_unique fails on it
					    * (sig/bip0340.c:428).  ECDSA keys are normalised on the device, where a key at infinity
					    * continues with W' = u*G like the reference's ec_verify. */
				memset(&pubs[i * 2 * plen], 0, 2 * plen);
			}
		});
	}
	if (sc == kBip0340) {
		/* the challenge hash needs x(Y) of the affine key: second marshalling pass */
		static const char tag[] = "BIP0340/challenge";
		uint8_t htag[128];
		const unsigned char *tin[2] = { (const unsigned char *)tag, nullptr };
		uint32_t tl[1] = { (uint32_t)(sizeof(tag) - 1) };
		if (hm->hfunc_scattered(tin, tl, htag)) return -1;
		parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
			for (uint32_t i = lo; i < hi; i++) {
				if (!ok[i]) continue;
				const unsigned char *in[6] = { htag, htag, s[i], &pubs[i * 2 * plen], m[i], nullptr };
				uint32_t il[5] = { hlen, hlen, (uint32_t)plen, (uint32_t)plen, m_len[i] };
				if (hm->hfunc_scattered(in, il, &dig[i * (size_t)hlen])) ok[i] = 0;
			}
		});
	}
	memset(verdict, 0xff, num);
	int rc = 0;
	/*
	 * ECFSDSA / BIP0340, large batches: first the whole batch as ONE multi-scalar multiplication (K6 — what the reference's
	 * own verify_batch computes with Bos-Coster, ~7x cheaper on the device than n individual verifications).  When it
	 * passes, every signature is valid; when it fails (or an item was already refused above, or s = 0, which the per-item
	 * ECFSDSA check excludes and the combination does not), the per-item kernel runs to say WHICH one is bad.
	 */
	bool settled = false;
	if ((sc == kEcfsdsa || sc == kBip0340) && num >= msm_min_batch()) {
		bool clean = true;
		for (uint32_t i = 0; i < num && clean; i++) clean = ok[i] != 0;
		if (clean && sc == kEcfsdsa) {
			std::atomic<int> zero_s{ 0 };
			parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
				for (uint32_t i = lo; i < hi; i++) {
					const uint8_t *sb = &sigs[i * siglen + 2 * plen];
					uint8_t acc = 0;
					for (size_t j = 0; j < qlen; j++) acc |= sb[j];
					if (!acc) zero_s.store(1);
				}
			});
			clean = zero_s.load() == 0;
		}
		if (clean) {
			int all_valid = 0;
			const int r = sc == kEcfsdsa
					      ? eccb200_ecfsdsa_verify_msm_batch(eng, num, sigs, pubs, dig, hlen, nullptr, &all_valid)
					      : eccb200_bip0340_verify_msm_batch(eng, num, sigs, pubs, dig, hlen, nullptr, &all_valid);
			if (r == 0 && all_valid == 1) { /* r != 0: not served on this curve (BIP0340 on SECP224R1) */
				memset(verdict, 0, num);
				settled = true;
				g_msm_batches += 1;
			}
		}
	}
	if (!settled) {
		if (sc == kEcfsdsa) rc = eccb200_ecfsdsa_verify_batch(eng, num, sigs, pubs, dig, hlen, verdict);
		else if (sc == kBip0340) rc = eccb200_bip0340_verify_batch(eng, num, sigs, pubs, dig, hlen, verdict);
		else rc = eccb200_ecdsa_verify_prj_batch(eng, num, sigs, pubs, dig, hlen, verdict);
	}
	if (rc) return -1;
	g_verifies += num;
	int all = 0;
	for (uint32_t i = 0; i < num; i++) {
		if (!ok[i]) verdict[i] = -1;
		if (verdict[i]) all = -1;
	}
	t_verdicts.assign(verdict, verdict + num);
	return all;
}

/*
 * Arithmetic modulo the group order q on the host, for the scalar preparations that multiply or invert (the
 * reference's nn_mod_mul / nn_modinv / nn_mod_add of sig/ecgdsa.c:563-569, sig/ecrdsa.c:556-570, sig/sm2.c:657,
 * sig/bign_common.c:906-916): Montgomery products on 64-bit limbs (R = 2^(64 n)), inverses by Fermat (q is prime),
 * shared among the items of a chunk with Montgomery's trick.  Values are little-endian limb arrays of n <= 9 words.
 */
struct ModQ {
	int n = 0;
	size_t qlen = 0;
	uint64_t q[9] = { 0 }, r2[9] = { 0 }, one[9] = { 0 }, n0 = 0;

	explicit ModQ(const CurveInfo *ci) : n(ci->n64), qlen((size_t)ci->qlen)
	{
		for (int i = 0; i < n; i++) q[i] = ci->q[i];
		uint64_t x = 1; /* -q^-1 mod 2^64 by Newton */
		for (int i = 0; i < 6; i++) x *= 2 - q[0] * x;
		n0 = 0 - x;
		uint64_t t[9] = { 1 }; /* 2^k mod q by doublings: k = 64n gives R, k = 128n gives R^2 */
		for (int k = 0; k < 128 * n; k++) {
			add(t, t, t);
			if (k == 64 * n - 1) memcpy(one, t, sizeof(one));
		}
		memcpy(r2, t, sizeof(r2));
	}
	bool is_zero(const uint64_t *a) const
	{
		uint64_t v = 0;
		for (int i = 0; i < n; i++) v |= a[i];
		return v == 0;
	}
	bool geq_q(const uint64_t *a) const
	{
		for (int i = n - 1; i >= 0; i--)
			if (a[i] != q[i]) return a[i] > q[i];
		return true;
	}
	bool eq(const uint64_t *a, const uint64_t *b) const { return memcmp(a, b, (size_t)n * 8) == 0; }
	void sub_q_if(uint64_t *a, uint64_t top) const /* a + top * 2^(64n) in [0, 2q) -> [0, q) */
	{
		if (!top && !geq_q(a)) return;
		unsigned __int128 bw = 0;
		for (int i = 0; i < n; i++) {
			unsigned __int128 d = (unsigned __int128)a[i] - q[i] - (uint64_t)bw;
			a[i] = (uint64_t)d;
			bw = (d >> 64) & 1;
		}
	}
	void add(uint64_t *o, const uint64_t *a, const uint64_t *b) const /* a, b < q */
	{
		unsigned __int128 c = 0;
		for (int i = 0; i < n; i++) {
			c += (unsigned __int128)a[i] + b[i];
			o[i] = (uint64_t)c;
			c >>= 64;
		}
		sub_q_if(o, (uint64_t)c);
	}
	void neg(uint64_t *o, const uint64_t *a) const /* a < q */
	{
		if (is_zero(a)) {
			memset(o, 0, (size_t)n * 8);
			return;
		}
		unsigned __int128 bw = 0;
		for (int i = 0; i < n; i++) {
			unsigned __int128 d = (unsigned __int128)q[i] - a[i] - (uint64_t)bw;
			o[i] = (uint64_t)d;
			bw = (d >> 64) & 1;
		}
	}
	/* o = a * b / R mod q; one operand < q, the other < R (o may alias a or b) */
	void mul(uint64_t *o, const uint64_t *a, const uint64_t *b) const
	{
		uint64_t t[11] = { 0 };
		for (int i = 0; i < n; i++) {
			unsigned __int128 c = 0;
			for (int j = 0; j < n; j++) {
				c += (unsigned __int128)a[i] * b[j] + t[j];
				t[j] = (uint64_t)c;
				c >>= 64;
			}
			c += t[n];
			t[n] = (uint64_t)c;
			t[n + 1] = (uint64_t)(c >> 64);
			const uint64_t m = t[0] * n0;
			c = ((unsigned __int128)m * q[0] + t[0]) >> 64;
			for (int j = 1; j < n; j++) {
				c += (unsigned __int128)m * q[j] + t[j];
				t[j - 1] = (uint64_t)c;
				c >>= 64;
			}
			c += t[n];
			t[n - 1] = (uint64_t)c;
			t[n] = t[n + 1] + (uint64_t)(c >> 64);
		}
		sub_q_if(t, t[n]);
		memcpy(o, t, (size_t)n * 8);
	}
	void to_mont(uint64_t *o, const uint64_t *a) const { mul(o, a, r2); }
	void from_mont(uint64_t *o, const uint64_t *a) const
	{
		const uint64_t u[9] = { 1 };
		mul(o, a, u);
	}
	void inv_mont(uint64_t *o, const uint64_t *a) const /* a^(q-2), Montgomery form in and out; a != 0 */
	{
		uint64_t e[9], acc[9], base[9];
		memcpy(e, q, sizeof(e));
		for (int i = 0, borrow = 2; i < n && borrow; i++) { /* e = q - 2 */
			const uint64_t old = e[i];
			e[i] = old - (uint64_t)borrow;
			borrow = old < (uint64_t)borrow ? 1 : 0;
		}
		memcpy(acc, one, sizeof(acc));
		memcpy(base, a, (size_t)n * 8);
		for (int i = 0; i < 64 * n; i++) {
			if ((e[i / 64] >> (i % 64)) & 1) mul(acc, acc, base);
			mul(base, base, base);
		}
		memcpy(o, acc, (size_t)n * 8);
	}
	/* simultaneous inversion of k nonzero plain values (v: k * 9 words, in place, plain in and out) */
	void inv_many(uint64_t *v, size_t k) const
	{
		if (!k) return;
		std::vector<uint64_t> pre(k * 9);
		uint64_t acc[9], t[9];
		memcpy(acc, one, sizeof(acc));
		for (size_t i = 0; i < k; i++) {
			to_mont(&v[i * 9], &v[i * 9]);
			memcpy(&pre[i * 9], acc, sizeof(acc)); /* product of the values before i */
			mul(acc, acc, &v[i * 9]);
		}
		inv_mont(acc, acc);
		for (size_t i = k; i-- > 0;) {
			mul(t, acc, &pre[i * 9]);      /* v_i^-1 (Montgomery form) */
			mul(acc, acc, &v[i * 9]);      /* drop v_i from the running inverse */
			from_mont(&v[i * 9], t);
		}
	}
	/* big-endian bytes of any length -> value mod q (nn_init_from_buf + nn_mod): Horner over 8n-byte chunks, a chunk
	 * (< R) reduced by a round trip through the Montgomery domain, h * R mod q = to_mont(h) */
	void from_be_mod(uint64_t *o, const uint8_t *be, size_t len) const
	{
		const size_t cb = (size_t)n * 8;
		uint64_t h[9] = { 0 };
		size_t pos = 0;
		size_t first = len % cb ? len % cb : (len ? cb : 0);
		while (pos < len) {
			const size_t take = pos == 0 ? first : cb;
			uint64_t c[9] = { 0 };
			for (size_t j = 0; j < take; j++) c[j / 8] |= (uint64_t)be[pos + take - 1 - j] << (8 * (j % 8));
			to_mont(c, c);
			from_mont(c, c);
			to_mont(h, h);
			add(h, h, c);
			pos += take;
		}
		memcpy(o, h, sizeof(h));
	}
	/* big-endian qlen bytes -> limbs, NOT reduced (for the range checks) */
	void from_be(uint64_t *o, const uint8_t *be) const
	{
		memset(o, 0, 9 * 8);
		for (size_t j = 0; j < qlen; j++) o[j / 8] |= (uint64_t)be[qlen - 1 - j] << (8 * (j % 8));
	}
	void to_be(uint8_t *out, const uint64_t *a) const { words_to_be(out, (int)qlen, a); }
	bool in_open_range(const uint64_t *a) const { return !is_zero(a) && !geq_q(a); } /* a in ]0, q[ */
};

/*
 * The verifications built on W' = a*G + b*Y (prj_pt_mul, prj_pt_mul, prj_pt_add, prj_pt_unique in the reference),
 * batched: ONE launch of eccb200_double_smul_batch per batch, the rest of the scheme on the host the way the reference
 * splits it — signature checks and the mod-q scalar preparation in front, the comparison (or the hash of the recomputed
 * point, src/hash) behind.  A scheme describes its three host steps; the skeleton (verify_batch_double_smul) does the
 * struct checks, brings the keys to affine form (on the device when they are projective), launches, collects verdicts.
 */
struct DsBatch {
	const CurveInfo *ci = nullptr;
	const HashMappingHead *hm = nullptr;
	uint32_t hlen = 0, num = 0;
	size_t plen = 0, qlen = 0;
	const uint8_t **s = nullptr;
	const uint8_t *s_len = nullptr;
	const eccb200_ec_pub_key **pub_keys = nullptr;
	const uint8_t **m = nullptr;
	const uint32_t *m_len = nullptr;
	const uint8_t **adata = nullptr;
	const uint16_t *adata_len = nullptr;
	uint8_t *ab = nullptr;   /* [num][2 qlen]  a || b, big-endian */
	uint8_t *pubs = nullptr; /* [num][2 plen]  affine keys, big-endian */
	std::vector<uint8_t> ok; /* item still alive */
	const ModQ *mq = nullptr;
	uint8_t *a_of(uint32_t i) const { return ab + (size_t)i * 2 * qlen; }
	uint8_t *b_of(uint32_t i) const { return ab + (size_t)i * 2 * qlen + qlen; }
	const uint8_t *key_of(uint32_t i) const { return pubs + (size_t)i * 2 * plen; }
	const uint8_t *id_of(uint32_t i) const { return adata ? adata[i] : nullptr; }
	uint16_t id_len_of(uint32_t i) const { return adata_len ? adata_len[i] : 0; }
};

struct DsScheme {
	virtual ~DsScheme() {}
	/* false: ancillary data makes the call one the layer does not serve (the caller forwards it) */
	virtual bool takes_adata() const { return false; }
	/* the scheme hashes the affine key (a key at infinity then fails prj_pt_to_aff in the reference); otherwise a
	 * key at infinity verifies like the reference: b * infinity = infinity, W' = a*G */
	virtual bool needs_affine_key() const { return false; }
	virtual bool setup(DsBatch &) { return true; }                          /* per batch: lengths, side buffers */
	virtual size_t siglen(const DsBatch &) const = 0;
	virtual bool sig_ok(const DsBatch &, uint32_t i) const = 0;             /* range checks of *_verify_init */
	virtual void scalars(DsBatch &, uint32_t lo, uint32_t hi) = 0;          /* fills a || b; may clear ok[i] */
	virtual bool accept(const DsBatch &, uint32_t i, const uint8_t *W) = 0; /* W = affine x || y of W', finite */
};

static bool be_in_open_range(const uint8_t *v, const uint8_t *qbe, size_t qlen) /* v in ]0, q[ */
{
	bool zero = true;
	for (size_t j = 0; j < qlen; j++) zero = zero && v[j] == 0;
	return !zero && memcmp(v, qbe, qlen) < 0;
}

/* ECSDSA / ECOSDSA (sig/ecsdsa_common.c:425-609): r = H(W_x [|| W_y] || m) is a digest, s a scalar;
 * e = -(OS2I(r) mod q); W' = sG + eY; accept iff H(W'_x [|| W'_y] || m) == r. */
struct EcsdsaScheme : DsScheme {
	bool optimized;
	uint8_t qbe[72];
	explicit EcsdsaScheme(bool opt) : optimized(opt) {}
	bool setup(DsBatch &b) override
	{
		words_to_be(qbe, (int)b.qlen, b.ci->q);
		return true;
	}
	size_t siglen(const DsBatch &b) const override { return (size_t)b.hlen + b.qlen; } /* ECSDSA_SIGLEN (:472) */
	bool sig_ok(const DsBatch &b, uint32_t i) const override
	{
		return be_in_open_range(b.s[i] + b.hlen, qbe, b.qlen);                      /* 1. s in ]0, q[ (:475-478) */
	}
	void scalars(DsBatch &b, uint32_t lo, uint32_t hi) override
	{
		for (uint32_t i = lo; i < hi; i++) {
			if (!b.ok[i]) continue;
			uint64_t r[9], e[9];
			b.mq->from_be_mod(r, b.s[i], b.hlen);                               /* 2. e = -(r mod q) (:486-488) */
			if (b.mq->is_zero(r)) {                                             /* 3. e == 0: reject (:491-492) */
				b.ok[i] = 0;
				continue;
			}
			b.mq->neg(e, r);
			memcpy(b.a_of(i), b.s[i] + b.hlen, b.qlen);                         /* 4. W' = sG + eY (:495-498) */
			b.mq->to_be(b.b_of(i), e);
		}
	}
	bool accept(const DsBatch &b, uint32_t i, const uint8_t *W) override
	{
		uint8_t rp[128];
		const unsigned char *in[4] = { W, optimized ? b.m[i] : W + b.plen, optimized ? nullptr : b.m[i], nullptr };
		uint32_t il[3] = { (uint32_t)b.plen, optimized ? b.m_len[i] : (uint32_t)b.plen, optimized ? 0u : b.m_len[i] };
		if (b.hm->hfunc_scattered(in, il, rp)) return false;                        /* 5. r' (:500-520) */
		return memcmp(rp, b.s[i], b.hlen) == 0;                                     /* 6. r == r' (:606) */
	}
};

/* ECKCDSA (sig/eckcdsa.c:543-832): r_len = min(|H|, qlen); s in ]0, q[; h = H(z || m) with z = the first block_size
 * bytes of Y_x || Y_y || 0...; e = OS2I(r XOR rightmost(h)) mod q; W' = sY + eG; r' = rightmost(H(W'_x)) == r. */
struct EckcdsaScheme : DsScheme {
	uint8_t qbe[72];
	size_t rlen = 0, shift = 0, zlen = 0;
	bool needs_affine_key() const override { return true; } /* z (:601-625): prj_pt_to_aff fails on infinity (:614) */
	bool setup(DsBatch &b) override
	{
		words_to_be(qbe, (int)b.qlen, b.ci->q);
		zlen = b.hm->block_size;
		rlen = std::min<size_t>(b.hlen, b.qlen); /* ECKCDSA_R_LEN (sig/eckcdsa.h:28-31) */
		shift = b.hlen > rlen ? b.hlen - rlen : 0;
		return zlen != 0;
	}
	size_t siglen(const DsBatch &b) const override { return rlen + b.qlen; }            /* 1. (:589) */
	bool sig_ok(const DsBatch &b, uint32_t i) const override
	{
		return be_in_open_range(b.s[i] + rlen, qbe, b.qlen);                         /* 2. s in ]0, q[ (:592-595) */
	}
	void scalars(DsBatch &b, uint32_t lo, uint32_t hi) override
	{
		std::vector<uint8_t> z(zlen);
		for (uint32_t i = lo; i < hi; i++) {
			if (!b.ok[i]) continue;
			std::fill(z.begin(), z.end(), 0);
			memcpy(z.data(), b.key_of(i), std::min<size_t>(zlen, 2 * b.plen));   /* 3. z (:601-625) */
			uint8_t h[128], x[128];
			const unsigned char *in[3] = { z.data(), b.m[i], nullptr };
			uint32_t il[2] = { (uint32_t)zlen, b.m_len[i] };
			if (b.hm->hfunc_scattered(in, il, h)) {
				b.ok[i] = 0;
				continue;
			}
			for (size_t j = 0; j < rlen; j++) x[j] = (uint8_t)(h[shift + j] ^ b.s[i][j]); /* 4.-5. (:754-762) */
			uint64_t e[9];
			b.mq->from_be_mod(e, x, rlen);
			b.mq->to_be(b.a_of(i), e);                                           /* e on G */
			memcpy(b.b_of(i), b.s[i] + rlen, b.qlen);                            /* s on Y: 6. W' = sY + eG (:770-773) */
		}
	}
	bool accept(const DsBatch &b, uint32_t i, const uint8_t *W) override
	{
		uint8_t rp[128];
		const unsigned char *in[2] = { W, nullptr };
		uint32_t il[1] = { (uint32_t)b.plen };
		if (b.hm->hfunc_scattered(in, il, rp)) return false;                         /* 7. r' = H(W'_x) (:778-784) */
		return memcmp(rp + shift, b.s[i], rlen) == 0;                                /* 8.-9. (:794-800) */
	}
};

/* x(W') mod q == r, r given as qlen big-endian bytes (nn_mod of the affine x, then nn_cmp) */
static bool x_mod_q_equals(const DsBatch &b, const uint8_t *W, const uint8_t *r_be)
{
	uint64_t x[9], r[9];
	b.mq->from_be_mod(x, W, b.plen);
	b.mq->from_be(r, r_be);
	return b.mq->eq(x, r);
}

/* leftmost min(8 hlen, bitlen(q)) bits of a digest as an integer mod q: the truncation ECDSA and ECGDSA share
 * (sig/ecgdsa.c:545-560: nn_rshift_fixedlen by 8 hlen - bitlen(q), nn_mod) */
static void digest_truncated_mod_q(const DsBatch &b, uint64_t *e, const uint8_t *h, int qbits)
{
	uint8_t buf[128];
	const int hbits = 8 * (int)b.hlen;
	const int rshift = hbits > qbits ? hbits - qbits : 0;
	/* h >> rshift, big-endian, still hlen bytes */
	const int bytes = rshift / 8, bits = rshift % 8;
	memset(buf, 0, sizeof(buf));
	for (int j = (int)b.hlen - 1; j >= bytes; j--) {
		unsigned v = h[j - bytes] >> bits;
		if (bits && j - bytes - 1 >= 0) v |= (unsigned)h[j - bytes - 1] << (8 - bits);
		buf[j] = (uint8_t)v;
	}
	b.mq->from_be_mod(e, buf, b.hlen);
}

static int order_bits(const CurveInfo *ci)
{
	for (int i = ci->n64 - 1; i >= 0; i--)
		if (ci->q[i]) return 64 * i + 64 - __builtin_clzll(ci->q[i]);
	return 0;
}

/* ECGDSA (sig/ecgdsa.c:413-600): r, s in ]0, q[; e = truncated H(m) mod q; u = r^-1 e, v = r^-1 s; W' = uG + vY;
 * accept iff W'_x mod q == r.  ECRDSA (sig/ecrdsa.c:417-600): s in ]0, q[, r != 0; h = OS2I(H(m)) mod q (the digest
 * byte-reversed unless the reference was built with USE_ISO14888_3_ECRDSA, :545-547), 0 replaced by 1; e = h^-1;
 * u = e s, v = -e r; the same W' and comparison.  One inversion mod q per chunk of items (Montgomery's trick). */
struct EcgdsaScheme : DsScheme {
	bool rdsa, iso_rdsa = false;
	int qbits = 0;
	explicit EcgdsaScheme(bool r) : rdsa(r) {}
	bool setup(DsBatch &b) override
	{
		qbits = order_bits(b.ci);
		const char *e = getenv("ECCB200_ECRDSA_ISO14888_3");
		iso_rdsa = e && atoi(e) != 0;
		return true;
	}
	size_t siglen(const DsBatch &b) const override { return 2 * b.qlen; } /* EC[GR]DSA_SIGLEN (ecgdsa.c:440, ecrdsa.c:443) */
	bool sig_ok(const DsBatch &b, uint32_t i) const override
	{
		uint64_t r[9], s[9];
		b.mq->from_be(r, b.s[i]);
		b.mq->from_be(s, b.s[i] + b.qlen);
		/* (ecgdsa.c:451-457).  ECRDSA's init compares s with q twice and never r (ecrdsa.c:453-456); an r >= q then
		 * fails at the final comparison with r' < q, so rejecting it here gives the same verdict. */
		return b.mq->in_open_range(r) && b.mq->in_open_range(s);
	}
	void scalars(DsBatch &b, uint32_t lo, uint32_t hi) override
	{
		const size_t cnt = hi - lo;
		std::vector<uint64_t> den(cnt * 9), e(cnt * 9);
		std::vector<uint32_t> idx;
		idx.reserve(cnt);
		for (uint32_t i = lo; i < hi; i++) {
			if (!b.ok[i]) continue;
			uint8_t h[128];
			const unsigned char *in[2] = { b.m[i], nullptr };
			uint32_t il[1] = { b.m_len[i] };
			if (b.hm->hfunc_scattered(in, il, h)) {                              /* 2. h = H(m) */
				b.ok[i] = 0;
				continue;
			}
			const size_t k = idx.size();
			if (rdsa) {
				if (!iso_rdsa) std::reverse(h, h + b.hlen);                  /* (ecrdsa.c:545-547) */
				b.mq->from_be_mod(&den[k * 9], h, b.hlen);                   /* 3. h mod q, 0 -> 1 (:550-555) */
				if (b.mq->is_zero(&den[k * 9])) den[k * 9] = 1;
			} else {
				digest_truncated_mod_q(b, &e[k * 9], h, qbits);              /* 3. e (ecgdsa.c:545-560) */
				b.mq->from_be(&den[k * 9], b.s[i]);                          /* r, to be inverted (:563) */
			}
			idx.push_back(i);
		}
		b.mq->inv_many(den.data(), idx.size());
		for (size_t k = 0; k < idx.size(); k++) {
			const uint32_t i = idx[k];
			uint64_t inv[9], r[9], s[9], u[9], v[9];
			b.mq->to_mont(inv, &den[k * 9]);
			b.mq->from_be(r, b.s[i]);
			b.mq->from_be(s, b.s[i] + b.qlen);
			if (rdsa) {
				b.mq->mul(u, inv, s);                                        /* 4. u = e s (ecrdsa.c:559) */
				b.mq->mul(v, inv, r);                                        /* 5. v = -e r (:568-569) */
				b.mq->neg(v, v);
			} else {
				b.mq->mul(u, inv, &e[k * 9]);                                /* 4. u = r^-1 e (ecgdsa.c:564) */
				b.mq->mul(v, inv, s);                                        /* 5. v = r^-1 s (:568) */
			}
			b.mq->to_be(b.a_of(i), u);                                           /* 6. W' = uG + vY */
			b.mq->to_be(b.b_of(i), v);
		}
	}
	bool accept(const DsBatch &b, uint32_t i, const uint8_t *W) override
	{
		return x_mod_q_equals(b, W, b.s[i]);                                         /* 7.-8. r' = W'_x mod q == r */
	}
};

/* SM2 (sig/sm2.c:518-700): r, s in ]0, q[; t = r + s mod q != 0; W' = sG + tY finite; Z = H(ENTL || ID || a || b ||
 * G_x || G_y || Y_x || Y_y) with the ID in the ancillary data (:140-200); e = OS2I(H(Z || m)) mod q;
 * accept iff (e + W'_x) mod q == r. */
struct Sm2Scheme : DsScheme {
	bool takes_adata() const override { return true; }
	bool needs_affine_key() const override { return true; } /* Z hashes it (prj_pt_export_to_aff_buf, :190) */
	size_t siglen(const DsBatch &b) const override { return 2 * b.qlen; }            /* SM2_SIGLEN (:546) */
	bool sig_ok(const DsBatch &b, uint32_t i) const override
	{
		uint64_t r[9], s[9];
		if (!b.id_of(i) || b.id_len_of(i) > 8191) return false;                      /* sm2_compute_Z (:149-151) */
		b.mq->from_be(r, b.s[i]);
		b.mq->from_be(s, b.s[i] + b.qlen);
		return b.mq->in_open_range(r) && b.mq->in_open_range(s);                     /* 1. (:553-557) */
	}
	void scalars(DsBatch &b, uint32_t lo, uint32_t hi) override
	{
		for (uint32_t i = lo; i < hi; i++) {
			if (!b.ok[i]) continue;
			uint64_t r[9], s[9], t[9];
			b.mq->from_be(r, b.s[i]);
			b.mq->from_be(s, b.s[i] + b.qlen);
			b.mq->add(t, r, s);                                                  /* 3. t = r + s mod q (:657) */
			if (b.mq->is_zero(t)) {                                              /* 4. (:660-661) */
				b.ok[i] = 0;
				continue;
			}
			memcpy(b.a_of(i), b.s[i] + b.qlen, b.qlen);                          /* 6. W' = sG + tY (:669-671) */
			b.mq->to_be(b.b_of(i), t);
		}
	}
	bool accept(const DsBatch &b, uint32_t i, const uint8_t *W) override
	{
		const eccb200_ec_shortw_crv *crv = b.pub_keys[i]->y.crv;
		if (!fp_ok(&crv->a) || !fp_ok(&crv->b)) return false;
		const int pl = (int)b.plen;
		uint8_t entl[2], ca[72], cb[72], g[144], Z[128], h[128];
		const uint16_t entlen = (uint16_t)(b.id_len_of(i) * 8);
		entl[0] = (uint8_t)(entlen >> 8);
		entl[1] = (uint8_t)entlen;
		fp_to_be(ca, &crv->a, pl);
		fp_to_be(cb, &crv->b, pl);
		gen_to_be(g, b.ci);
		const unsigned char *zin[7] = { entl, b.id_of(i), ca, cb, g, b.key_of(i), nullptr };
		uint32_t zl[6] = { 2, b.id_len_of(i), (uint32_t)pl, (uint32_t)pl, (uint32_t)(2 * pl), (uint32_t)(2 * pl) };
		if (b.hm->hfunc_scattered(zin, zl, Z)) return false;                         /* Z (:160-196) */
		const unsigned char *in[3] = { Z, b.m[i], nullptr };
		uint32_t il[2] = { b.hlen, b.m_len[i] };
		if (b.hm->hfunc_scattered(in, il, h)) return false;                          /* 2. h = H(Z || m) */
		uint64_t e[9], x[9], r[9];
		b.mq->from_be_mod(e, h, b.hlen);                                             /* 5. e (:664-666) */
		b.mq->from_be_mod(x, W, b.plen);                                             /* 8. r' = (e + W'_x) mod q (:679-684) */
		b.mq->add(x, x, e);
		b.mq->from_be(r, b.s[i]);
		return b.mq->eq(x, r);                                                       /* 9. (:687-688) */
	}
};

/* BIGN / DBIGN (sig/bign_common.c:742-990), little-endian byte strings: signature s0 (l = qlen / 2 bytes) || s1 (qlen
 * bytes), s1 < q; h = OS2I_le(H(m)) mod q; W = ((s1 + h) mod q) G + ((s0 + 2^(8l)) mod q) Y finite;
 * t = first l bytes of BELT-HASH(OID || first 2l bytes of LE(W_x) || LE(W_y) || H(m)), the OID in the ancillary data
 * (:97-121); accept iff t == s0. */
struct BignScheme : DsScheme {
	size_t l = 0;
	const HashMappingHead *belt = nullptr;
	std::vector<uint8_t> digests; /* H(m_i), needed again behind the launch */
	bool takes_adata() const override { return true; }
	bool setup(DsBatch &b) override
	{
		l = b.qlen / 2; /* BIGN_S0_LEN (sig/bign_common.h:34) */
		get_hash_fn get_hash = resolve_get_hash();
		if (!get_hash || get_hash(16 /* BELT_HASH, lib_ecc_types.h:130 */, &belt) || !belt || !belt->hfunc_scattered) return false;
		if (belt->digest_size > 64) return false;
		digests.assign((size_t)b.num * b.hlen, 0);
		return true;
	}
	size_t siglen(const DsBatch &b) const override { return l + b.qlen; }            /* BIGN_SIGLEN (:775) */
	static bool oid_of(const uint8_t *adata, uint16_t adata_len, const uint8_t **oid, uint16_t *oid_len)
	{
		if (!adata || adata_len < 4) return false;                                   /* bign_get_oid_from_adata (:97-112) */
		const uint32_t ol = ((uint32_t)adata[0] << 8) | adata[1], tl = ((uint32_t)adata[2] << 8) | adata[3];
		if (ol + tl > (uint32_t)(adata_len - 4)) return false; /* the reference's sum is an int: no wrap */
		*oid = adata + 4;
		*oid_len = (uint16_t)ol;
		return true;
	}
	bool sig_ok(const DsBatch &b, uint32_t i) const override
	{
		const uint8_t *oid;
		uint16_t ol;
		if (!b.id_of(i) || b.id_len_of(i) == 0) return false;                        /* (:763) */
		if (!oid_of(b.id_of(i), b.id_len_of(i), &oid, &ol)) return false;            /* fails at finalize (:929) */
		uint8_t be[72];
		uint64_t s1[9];
		for (size_t j = 0; j < b.qlen; j++) be[j] = b.s[i][l + b.qlen - 1 - j];
		b.mq->from_be(s1, be);
		return !b.mq->geq_q(s1);                                                     /* 1. s1 < q (:790-791) */
	}
	void scalars(DsBatch &b, uint32_t lo, uint32_t hi) override
	{
		uint64_t two_l[9] = { 0 }; /* 2^(8l) mod q (:909-912): 8l < bitlen(q) whenever qlen >= 2 */
		{
			uint8_t be[72] = { 0 };
			be[0] = 1;
			b.mq->from_be_mod(two_l, be, l + 1);
		}
		for (uint32_t i = lo; i < hi; i++) {
			if (!b.ok[i]) continue;
			uint8_t *hd = &digests[(size_t)i * b.hlen], hr[128], be[72];
			const unsigned char *in[2] = { b.m[i], nullptr };
			uint32_t il[1] = { b.m_len[i] };
			if (b.hm->hfunc_scattered(in, il, hd)) {                             /* 2. h = H(m) */
				b.ok[i] = 0;
				continue;
			}
			std::reverse_copy(hd, hd + b.hlen, hr);                              /* (:898-900) */
			uint64_t h[9], s1[9], s0[9], u[9], v[9];
			b.mq->from_be_mod(h, hr, b.hlen);
			for (size_t j = 0; j < b.qlen; j++) be[j] = b.s[i][l + b.qlen - 1 - j];
			b.mq->from_be(s1, be);
			b.mq->add(u, h, s1);                                                 /* (s1 + h) mod q (:906) */
			for (size_t j = 0; j < l; j++) be[j] = b.s[i][l - 1 - j];
			b.mq->from_be_mod(s0, be, l);
			b.mq->add(v, two_l, s0);                                             /* (s0 + 2^(8l)) mod q (:909-913) */
			b.mq->to_be(b.a_of(i), u);                                           /* 3. W (:916-918) */
			b.mq->to_be(b.b_of(i), v);
		}
	}
	bool accept(const DsBatch &b, uint32_t i, const uint8_t *W) override
	{
		const uint8_t *oid = nullptr;
		uint16_t ol = 0;
		if (!oid_of(b.id_of(i), b.id_len_of(i), &oid, &ol)) return false;
		uint8_t le[144], hb[64], t[72];
		for (size_t j = 0; j < b.plen; j++) {                                        /* FE2OS(W_x) || FE2OS(W_y) (:932-936) */
			le[j] = W[b.plen - 1 - j];
			le[b.plen + j] = W[2 * b.plen - 1 - j];
		}
		const unsigned char *in[4] = { oid, le, &digests[(size_t)i * b.hlen], nullptr };
		uint32_t il[3] = { ol, (uint32_t)(2 * l), b.hlen };
		if (belt->hfunc_scattered(in, il, hb)) return false;                         /* 6. (:927-944) */
		memset(t, 0, sizeof(t));
		memcpy(t, hb, std::min<size_t>(l, belt->digest_size));                       /* (:946-947) */
		return memcmp(t, b.s[i], l) == 0;                                            /* 10. t == s0 (:950-951) */
	}
};

/* ECCB200_DROPIN_TIMING=1: wall-clock time of the phases of a batch on stderr (where the host side of an adapter goes) */
struct PhaseClock {
	bool on;
	std::chrono::steady_clock::time_point t0;
	std::string line;
	PhaseClock() : on(getenv("ECCB200_DROPIN_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
	void lap(const char *name)
	{
		if (!on) return;
		const auto t1 = std::chrono::steady_clock::now();
		char buf[64];
		snprintf(buf, sizeof buf, " %s %.2f ms", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
		line += buf;
		t0 = t1;
	}
	void report(uint32_t n) const
	{
		if (on) fprintf(stderr, "dropin timing (%u items):%s\n", n, line.c_str());
	}
};

static int verify_batch_double_smul(DsScheme &sch, const uint8_t **s, const uint8_t *s_len,
				    const eccb200_ec_pub_key **pub_keys, const uint8_t **m, const uint32_t *m_len,
				    uint32_t num, int sig_type, int hash_type, const uint8_t **adata,
				    const uint16_t *adata_len)
{
	if (num > kMaxBatch) {
		t_verdicts.clear();
		return -1;
	}
	t_verdicts.assign(num, -1);
	if (num == 0) return -1;
	if (!s || !s_len || !pub_keys || !m || !m_len) return -1;
	if (adata && !sch.takes_adata())
		for (uint32_t i = 0; i < num; i++)
			if (adata[i]) return -1;
	get_hash_fn get_hash = resolve_get_hash();
	if (!get_hash) return -1;
	DsBatch b;
	if (get_hash(hash_type, &b.hm) || !b.hm || !b.hm->hfunc_scattered) return -1;
	b.hlen = b.hm->digest_size;
	if (b.hlen == 0 || b.hlen > 128) return -1;
	for (uint32_t i = 0; i < num && !b.ci; i++) {
		const eccb200_ec_pub_key *pk = pub_keys[i];
		if (pk && pk->magic == kPubKeyMagic && pk->key_type == sig_type && pt_ok(&pk->y)) b.ci = identify(&pk->y);
	}
	if (!b.ci) return -1;
	const CurveInfo *ci = b.ci;
	Engine engine = acquire(ci->id, num);
	eccb200_ctx *eng = engine.ctx;
	if (!eng) return -1;
	const ModQ mq(ci);
	const int pl = ci->plen;
	const size_t plen = (size_t)ci->plen, qlen = (size_t)ci->qlen;
	b.num = num;
	b.plen = plen;
	b.qlen = qlen;
	b.s = s;
	b.s_len = s_len;
	b.pub_keys = pub_keys;
	b.m = m;
	b.m_len = m_len;
	b.adata = adata;
	b.adata_len = adata_len;
	b.mq = &mq;
	b.ab = engine.slot->st[0].get(num * 2 * qlen);
	b.pubs = engine.slot->st[1].get(num * 2 * plen);
	uint8_t *wout = engine.slot->st[2].get(num * 2 * plen);
	int8_t *status = (int8_t *)engine.slot->st[3].get(num);
	if (!b.ab || !b.pubs || !wout || !status) return -1;
	if (!sch.setup(b)) return -1;
	const size_t siglen = sch.siglen(b);
	b.ok.assign(num, 0);
	std::vector<uint8_t> key_inf(num, 0);
	std::atomic<int> mixed{ 0 }, any_prj{ 0 };
	/* an ec_pub_key made by ec_key_pair_gen holds a projective point with Z != 1: X || Y || Z of every item goes into
	 * page-locked staging and, if any key needs it, through ONE batched prj_pt_unique on the device, which writes the
	 * affine keys straight into the staging the double-scalar launch reads (as verify_batch_common does) */
	uint8_t *prj = engine.slot->st[4].get(num * 3 * plen);
	int8_t *prj_st = (int8_t *)engine.slot->st[5].get(num);
	if (!prj || !prj_st) return -1;
	PhaseClock clk;
	/* pass 1: struct and signature checks, key marshalling */
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
		bool saw_prj = false;
		for (uint32_t i = lo; i < hi; i++) {
			memset(b.a_of(i), 0, 2 * qlen);
			memset(&b.pubs[i * 2 * plen], 0, 2 * plen);
			memset(&prj[i * 3 * plen], 0, 3 * plen);
			const eccb200_ec_pub_key *pk = pub_keys[i];
			if (!pk || pk->magic != kPubKeyMagic || pk->key_type != sig_type || !pt_ok(&pk->y)) continue;
			if (!s[i] || (!m[i] && m_len[i])) continue;
			const CurveInfo *c = identify(&pk->y);
			if (!c) continue;
			if (c != ci) {
				mixed.store(1);
				continue;
			}
			if (s_len[i] != siglen) continue;
			if (!sch.sig_ok(b, i)) continue;
			const eccb200_prj_pt *y = &pk->y;
			const bool affine = fp_is_small(&y->Z, 1);
			if (affine) {
				fp_to_be(&b.pubs[i * 2 * plen], &y->X, pl);
				fp_to_be(&b.pubs[i * 2 * plen + plen], &y->Y, pl);
			}
			saw_prj = saw_prj || !affine;
			fp_to_be(&prj[i * 3 * plen], &y->X, pl);
			fp_to_be(&prj[i * 3 * plen + plen], &y->Y, pl);
			fp_to_be(&prj[i * 3 * plen + 2 * plen], &y->Z, pl);
			b.ok[i] = 1;
		}
		if (saw_prj) any_prj.store(1);
	});
	if (mixed.load()) return -1;
	clk.lap("checks");
	if (any_prj.load()) {
		/* every item's key through the batched prj_pt_unique (items refused above carry Z = 0 and are ignored) */
		if (eccb200_prj_pt_unique_batch(eng, num, prj, b.pubs, prj_st)) return -1;
		parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
			for (uint32_t i = lo; i < hi; i++) {
				if (!b.ok[i] || prj_st[i] == 0) continue;
				if (prj_st[i] == 1 && !sch.needs_affine_key()) key_inf[i] = 1; /* b * infinity = infinity: W' = a*G */
				else b.ok[i] = 0; /* off the curve; or infinity where the scheme exports the affine key */
			}
		});
	}
	clk.lap("keys");
	/* pass 2: the scheme's scalars a, b */
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) { sch.scalars(b, lo, hi); });
	clk.lap("scalars");
	for (uint32_t i = 0; i < num; i++) {
		if (!b.ok[i]) { /* keep the batch launchable: rejected slots multiply the generator by zero */
			memset(b.a_of(i), 0, 2 * qlen);
			gen_to_be(&b.pubs[i * 2 * plen], ci);
		} else if (key_inf[i]) {
			memset(b.b_of(i), 0, qlen);
			gen_to_be(&b.pubs[i * 2 * plen], ci);
		}
	}
	if (eccb200_double_smul_batch(eng, num, b.ab, b.pubs, wout, status)) return -1;
	clk.lap("device");
	g_verifies += num;
	/* pass 3: the scheme's acceptance test on the affine W' */
	std::vector<int8_t> verdict(num, -1);
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
		for (uint32_t i = lo; i < hi; i++) {
			if (!b.ok[i] || status[i] != 0) continue; /* W' at infinity: prj_pt_unique / the explicit test fails */
			verdict[i] = sch.accept(b, i, &wout[i * 2 * plen]) ? 0 : -1;
		}
	});
	int all = 0;
	for (uint32_t i = 0; i < num; i++)
		if (verdict[i]) all = -1;
	t_verdicts.assign(verdict.begin(), verdict.end());
	clk.lap("accept");
	clk.report(num);
	return all;
}

/* the scheme object for an ec_alg_type served through verify_batch_double_smul (lib_ecc_types.h:22-80), or null */
static std::unique_ptr<DsScheme> ds_scheme_of(int sig_type)
{
	switch (sig_type) {
	case 2: return std::unique_ptr<DsScheme>(new EckcdsaScheme());
	case 3: return std::unique_ptr<DsScheme>(new EcsdsaScheme(false));
	case 4: return std::unique_ptr<DsScheme>(new EcsdsaScheme(true)); /* ECOSDSA */
	case 6: return std::unique_ptr<DsScheme>(new EcgdsaScheme(false));
	case 7: return std::unique_ptr<DsScheme>(new EcgdsaScheme(true)); /* ECRDSA */
	case 8: return std::unique_ptr<DsScheme>(new Sm2Scheme());
	case 18:
	case 19: return std::unique_ptr<DsScheme>(new BignScheme()); /* BIGN, DBIGN (deterministic nonce: same verification) */
	default: return nullptr;
	}
}

static int verify_batch_ds(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
			   const uint32_t *m_len, uint32_t num, int sig_type, int hash_type, const uint8_t **adata,
			   const uint16_t *adata_len)
{
	std::unique_ptr<DsScheme> sch = ds_scheme_of(sig_type);
	if (!sch) {
		t_verdicts.assign(num, -1);
		return -1;
	}
	return verify_batch_double_smul(*sch, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata, adata_len);
}

extern "C" int eccb200_dropin_ecdsa_verify_batch(const uint8_t **s, const uint8_t *s_len,
						 const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
						 const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
						 const uint8_t **adata, const uint16_t *adata_len,
						 void *scratch_pad_area, uint32_t *scratch_pad_area_len)
{
	(void)scratch_pad_area;
	(void)scratch_pad_area_len;
	(void)adata_len;
	if (sig_type != 1 /* ECDSA */ && sig_type != 14 /* DECDSA */) return -1;
	return verify_batch_common(kEcdsa, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
}

/* ECFSDSA: a replacement for the reference's own ecfsdsa_verify_batch (sig/ecfsdsa.c:1057) in the same slot */
extern "C" int eccb200_dropin_ecfsdsa_verify_batch(const uint8_t **s, const uint8_t *s_len,
						   const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
						   const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
						   const uint8_t **adata, const uint16_t *adata_len,
						   void *scratch_pad_area, uint32_t *scratch_pad_area_len)
{
	(void)scratch_pad_area;
	(void)scratch_pad_area_len;
	(void)adata_len;
	if (sig_type != 5 /* ECFSDSA */) return -1;
	return verify_batch_common(kEcfsdsa, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
}

/* ECSDSA (sig_type 3) / ECOSDSA (4): the slot these two leave at unsupported_verify_batch in the reference */
extern "C" int eccb200_dropin_ecsdsa_verify_batch(const uint8_t **s, const uint8_t *s_len,
						  const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
						  const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
						  const uint8_t **adata, const uint16_t *adata_len,
						  void *scratch_pad_area, uint32_t *scratch_pad_area_len)
{
	(void)scratch_pad_area;
	(void)scratch_pad_area_len;
	if (sig_type != 3 /* ECSDSA */ && sig_type != 4 /* ECOSDSA */) return -1;
	return verify_batch_ds(s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata, adata_len);
}

/* ECKCDSA (sig_type 2): also left at unsupported_verify_batch by the reference */
extern "C" int eccb200_dropin_eckcdsa_verify_batch(const uint8_t **s, const uint8_t *s_len,
						   const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
						   const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
						   const uint8_t **adata, const uint16_t *adata_len,
						   void *scratch_pad_area, uint32_t *scratch_pad_area_len)
{
	(void)scratch_pad_area;
	(void)scratch_pad_area_len;
	if (sig_type != 2 /* ECKCDSA */) return -1;
	return verify_batch_ds(s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata, adata_len);
}

/*
 * ECGDSA (6), ECRDSA (7), SM2 (8), BIGN (18) / DBIGN (19): all left at unsupported_verify_batch by the reference
 * (sig/sig_algs_internal.h).  Their EC core W' = a*G + b*Y is one device launch for the batch; the mod-q scalar
 * preparation (with ONE inversion per chunk of items where the scheme inverts), the hashes and the comparison stay on the
 * host.  SM2 and BIGN read their ancillary data (the signer's ID; the hash OID) from adata[i] / adata_len[i].
 */
#define ECCB200_DS_ADAPTER(name, cond)                                                                                  \
	extern "C" int name(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys, const uint8_t **m, \
			    const uint32_t *m_len, uint32_t num, int sig_type, int hash_type, const uint8_t **adata,      \
			    const uint16_t *adata_len, void *scratch_pad_area, uint32_t *scratch_pad_area_len)            \
	{                                                                                                               \
		(void)scratch_pad_area;                                                                                 \
		(void)scratch_pad_area_len;                                                                             \
		if (!(cond)) return -1;                                                                                 \
		return verify_batch_ds(s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata, adata_len);       \
	}
ECCB200_DS_ADAPTER(eccb200_dropin_ecgdsa_verify_batch, sig_type == 6)
ECCB200_DS_ADAPTER(eccb200_dropin_ecrdsa_verify_batch, sig_type == 7)
ECCB200_DS_ADAPTER(eccb200_dropin_sm2_verify_batch, sig_type == 8)
ECCB200_DS_ADAPTER(eccb200_dropin_bign_verify_batch, sig_type == 18 || sig_type == 19)
#undef ECCB200_DS_ADAPTER

/* BIP0340: a replacement for the reference's own bip0340_verify_batch (sig/bip0340.c:1296) in the same slot */
extern "C" int eccb200_dropin_bip0340_verify_batch(const uint8_t **s, const uint8_t *s_len,
						   const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
						   const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
						   const uint8_t **adata, const uint16_t *adata_len,
						   void *scratch_pad_area, uint32_t *scratch_pad_area_len)
{
	(void)scratch_pad_area;
	(void)scratch_pad_area_len;
	(void)adata_len;
	if (sig_type != 20 /* BIP0340 */) return -1;
	return verify_batch_common(kBip0340, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
}

/*
 * ec_verify with the reference's exact prototype (sig/sig_algs.h:85-88, sig/sig_algs.c:655).  ECDSA / DECDSA and
 * ECFSDSA without ancillary data are verified by ONE launch of the verification kernel (hash on the host with the
 * reference's src/hash, then u*G + v*Y and the comparison on the device) instead of the reference's host code with
 * two interposed prj_pt_mul round trips; every other scheme — and anything this layer cannot serve (unknown curve,
 * no hash mapping) — is forwarded to the next ec_verify in the process, i.e. the reference's own.
 */
typedef int (*ec_verify_sig)(const uint8_t *, uint8_t, const eccb200_ec_pub_key *, const uint8_t *, uint32_t, int, int,
			     const uint8_t *, uint16_t);

extern "C" int ec_verify(const uint8_t *sig, uint8_t siglen, const eccb200_ec_pub_key *pub_key, const uint8_t *m,
			 uint32_t mlen, int sig_type, int hash_type, const uint8_t *adata, uint16_t adata_len);

static ec_verify_sig next_ec_verify()
{
	static ec_verify_sig next = (ec_verify_sig)next_definition("ec_verify");
	return next;
}

extern "C" int eccb200_dropin_ec_verify(const uint8_t *sig, uint8_t siglen, const eccb200_ec_pub_key *pub_key,
					const uint8_t *m, uint32_t mlen, int sig_type, int hash_type,
					const uint8_t *adata, uint16_t adata_len)
{
	Scheme sc = kEcdsa;
	std::unique_ptr<DsScheme> ds = ds_scheme_of(sig_type); /* the schemes built on W' = a*G + b*Y with host steps around */
	/* ancillary data: SM2 and BIGN need theirs (a call without is left to the reference, which rejects it); a call
	 * of any other scheme that carries some is forwarded untouched */
	const bool adata_fits = ds && ds->takes_adata() ? adata != nullptr : (!adata && adata_len == 0);
	bool ours = (ds || scheme_of(sig_type, &sc)) && adata_fits && sig && pub_key &&
		    pub_key->magic == kPubKeyMagic &&
		    pub_key->key_type == sig_type && pt_ok(&pub_key->y) && identify(&pub_key->y) != nullptr &&
		    resolve_get_hash() != nullptr;
	if (ours) {
		const HashMappingHead *hm = nullptr;
		ours = !resolve_get_hash()(hash_type, &hm) && hm && hm->hfunc_scattered;
	}
	if (!ours) {
		ec_verify_sig next = next_ec_verify();
		return next ? next(sig, siglen, pub_key, m, mlen, sig_type, hash_type, adata, adata_len) : -1;
	}
	const uint8_t *sp[1] = { sig };
	const uint8_t sl[1] = { siglen };
	const eccb200_ec_pub_key *pk[1] = { pub_key };
	const uint8_t *mp[1] = { m };
	const uint32_t ml[1] = { mlen };
	const uint8_t *ap[1] = { adata };
	const uint16_t al[1] = { adata_len };
	if (ds) return verify_batch_double_smul(*ds, sp, sl, pk, mp, ml, 1, sig_type, hash_type, adata ? ap : nullptr, al);
	return verify_batch_common(sc, sp, sl, pk, mp, ml, 1, sig_type, hash_type, nullptr);
}

extern "C" int ec_verify(const uint8_t *sig, uint8_t siglen, const eccb200_ec_pub_key *pub_key, const uint8_t *m,
			 uint32_t mlen, int sig_type, int hash_type, const uint8_t *adata, uint16_t adata_len)
{
	return eccb200_dropin_ec_verify(sig, siglen, pub_key, m, mlen, sig_type, hash_type, adata, adata_len);
}

/*
 * ec_verify_batch and is_verify_batch_mode_supported with the reference's exact prototypes (sig/sig_algs.h:90-93,
 * sig/sig_algs.c:675-694 and :937-958).  The reference dispatches through ec_sig_maps[].verify_batch, where ECDSA,
 * DECDSA, ECSDSA, ECOSDSA and ECKCDSA sit at unsupported_verify_batch (sig/sig_algs_internal.h:294); here those five
 * and ECFSDSA / BIP0340 are served by the device; every other scheme, an unknown curve or a batch with ancillary data
 * goes to the next definition in the process (the reference's own), -1 if there is none.
 */
typedef int (*ec_verify_batch_sig)(const uint8_t **, const uint8_t *, const eccb200_ec_pub_key **, const uint8_t **,
				   const uint32_t *, uint32_t, int, int, const uint8_t **, const uint16_t *, void *,
				   uint32_t *);
typedef int (*batch_supported_sig)(int, int *);

static bool batch_scheme_served(int sig_type)
{
	Scheme sc;
	return scheme_of(sig_type, &sc) || ds_scheme_of(sig_type) != nullptr;
}

extern "C" int eccb200_dropin_ec_verify_batch(const uint8_t **s, const uint8_t *s_len,
					      const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
					      const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
					      const uint8_t **adata, const uint16_t *adata_len, void *scratch_pad_area,
					      uint32_t *scratch_pad_area_len)
{
	std::unique_ptr<DsScheme> ds = ds_scheme_of(sig_type);
	bool ours = batch_scheme_served(sig_type) && num > 0 && s && s_len && pub_keys && m && m_len &&
		    resolve_get_hash() != nullptr;
	if (ours && adata && !(ds && ds->takes_adata())) /* ancillary data on a scheme that has none: not ours */
		for (uint32_t i = 0; i < num && ours; i++) ours = adata[i] == nullptr;
	if (ours) { /* a curve this layer knows, named by the first well-formed key */
		const CurveInfo *ci = nullptr;
		for (uint32_t i = 0; i < num && !ci; i++) {
			const eccb200_ec_pub_key *pk = pub_keys[i];
			if (pk && pk->magic == kPubKeyMagic && pk->key_type == sig_type && pt_ok(&pk->y))
				ci = identify(&pk->y);
		}
		const HashMappingHead *hm = nullptr;
		ours = ci != nullptr && !resolve_get_hash()(hash_type, &hm) && hm && hm->hfunc_scattered;
	}
	if (!ours) {
		static ec_verify_batch_sig next = (ec_verify_batch_sig)next_definition("ec_verify_batch");
		return next ? next(s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata, adata_len,
				   scratch_pad_area, scratch_pad_area_len)
			    : -1;
	}
	Scheme sc = kEcdsa;
	if (ds) return verify_batch_double_smul(*ds, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata, adata_len);
	scheme_of(sig_type, &sc);
	return verify_batch_common(sc, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
}

extern "C" int ec_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
			       const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
			       const uint8_t **adata, const uint16_t *adata_len, void *scratch_pad_area,
			       uint32_t *scratch_pad_area_len)
{
	return eccb200_dropin_ec_verify_batch(s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata, adata_len,
					      scratch_pad_area, scratch_pad_area_len);
}

extern "C" int is_verify_batch_mode_supported(int sig_type, int *check)
{
	if (!check) return -1;
	if (batch_scheme_served(sig_type)) {
		*check = 1;
		return 0;
	}
	static batch_supported_sig next = (batch_supported_sig)next_definition("is_verify_batch_mode_supported");
	return next ? next(sig_type, check) : -1;
}
