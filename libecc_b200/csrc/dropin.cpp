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
#include <cstdlib>
#include <cstring>
#include <functional>
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
	unsigned hw = std::thread::hardware_concurrency();
	if (const char *e = getenv("ECCB200_DROPIN_THREADS")) hw = (unsigned)atoi(e);
	unsigned T = std::min<unsigned>(std::min<unsigned>(hw ? hw : 1, 32u), n / 2048u + 1u);
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
	if (!out || !m || !in) return -1;
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
	Engine engine = acquire(ci->id, n);
	eccb200_ctx *eng = engine.ctx;
	if (!eng) return -1;
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
		if (eccb200_prj_pt_unique_batch(eng, (uint32_t)prj_idx.size(), pb.data(), ab.data(), st.data())) return -1;
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
		return -1;
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
			for (uint32_t i = lo; i < hi; i++)
				if (ok[i] && prj_st[i] != 0) {
					ok[i] = 0; /* off the curve, or the point at infinity: BIP0340's prj_pt_unique fails on it
						    * (sig/bip0340.c:428); ECFSDSA with a key at infinity is rejected too — a documented
						    * divergence (the reference would accept it iff s*G == r).  ECDSA keys are
						    * normalised on the device, where a key at infinity continues with W' = u*G like the
						    * reference's ec_verify. */
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

/* big-endian bytes -> nn-style little-endian 64-bit words (at most kMaxWords) */
static void be_to_nn(eccb200_nn *out, const uint8_t *be, uint32_t len)
{
	memset(out, 0, sizeof(*out));
	for (uint32_t j = 0; j < len && j < 8u * kMaxWords; j++) out->val[j / 8] |= (uint64_t)be[len - 1 - j] << (8 * (j % 8));
	out->wlen = (uint8_t)((std::min<uint32_t>(len, 8u * kMaxWords) + 7) / 8);
	out->magic = kNnMagic;
}

/*
 * ECSDSA / ECOSDSA (sig/ecsdsa_common.c:425-609): these schemes hash the RECOMPUTED point, so the batch is split the way
 * the reference's own code is: host — checks, e = -(OS2I(r) mod q); device — W' = sG + eY for the whole batch in one
 * launch (eccb200_double_smul_batch); host — r' = H(W'x [|| W'y] || m) with the reference's src/hash, r' == r.
 */
static int verify_batch_ecsdsa(bool optimized, const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
			       const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
			       const uint8_t **adata)
{
	t_verdicts.assign(num, -1);
	if (num == 0) return -1;
	if (!s || !s_len || !pub_keys || !m || !m_len) return -1;
	if (adata)
		for (uint32_t i = 0; i < num; i++)
			if (adata[i]) return -1;
	get_hash_fn get_hash = resolve_get_hash();
	if (!get_hash) return -1;
	const HashMappingHead *hm = nullptr;
	if (get_hash(hash_type, &hm) || !hm || !hm->hfunc_scattered) return -1;
	const uint32_t hlen = hm->digest_size;
	if (hlen == 0 || hlen > 128) return -1;
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
	const size_t siglen = (size_t)hlen + qlen; /* ECSDSA_SIGLEN: r is a digest, s a scalar (sig/ecsdsa_common.h) */
	uint8_t *ab = engine.slot->st[0].get(num * 2 * qlen), *pubs = engine.slot->st[1].get(num * 2 * plen),
		*wout = engine.slot->st[2].get(num * 2 * plen);
	int8_t *status = (int8_t *)engine.slot->st[3].get(num);
	if (!ab || !pubs || !wout || !status) return -1;
	uint8_t qbe[72];
	words_to_be(qbe, (int)qlen, ci->q);
	std::vector<uint8_t> ok(num, 0);
	std::atomic<int> mixed{ 0 };
	std::vector<std::vector<uint32_t>> prj_parts(64);
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned t) {
		std::vector<uint32_t> &prj = prj_parts[t];
		for (uint32_t i = lo; i < hi; i++) {
			memset(&ab[i * 2 * qlen], 0, 2 * qlen);
			memset(&pubs[i * 2 * plen], 0, 2 * plen);
			const eccb200_ec_pub_key *pk = pub_keys[i];
			if (!pk || pk->magic != kPubKeyMagic || pk->key_type != sig_type || !pt_ok(&pk->y)) continue;
			if (!s[i] || (!m[i] && m_len[i])) continue;
			const CurveInfo *c = identify(&pk->y);
			if (!c) continue;
			if (c != ci) {
				mixed.store(1);
				continue;
			}
			if (s_len[i] != siglen) continue;                       /* (:472) */
			const uint8_t *sb = s[i] + hlen;
			bool zero = true;
			for (size_t j = 0; j < qlen; j++) zero = zero && sb[j] == 0;
			if (zero || memcmp(sb, qbe, qlen) >= 0) continue;       /* 1. s in ]0, q[ (:475-478) */
			eccb200_nn r;
			be_to_nn(&r, s[i], hlen);
			uint8_t rmod[72], e[72];
			scalar_mod_to_be(rmod, &r, ci);                         /* 2. e = -(r mod q) mod q (:486-488) */
			bool rz = true;
			for (size_t j = 0; j < qlen; j++) rz = rz && rmod[j] == 0;
			if (rz) continue;                                       /* 3. e == 0: reject (:491-492) */
			int borrow = 0;
			for (int j = (int)qlen - 1; j >= 0; j--) {
				int d = (int)qbe[j] - (int)rmod[j] - borrow;
				borrow = d < 0;
				e[j] = (uint8_t)(d + (borrow << 8));
			}
			memcpy(&ab[i * 2 * qlen], sb, qlen);
			memcpy(&ab[i * 2 * qlen + qlen], e, qlen);
			const eccb200_prj_pt *y = &pk->y;
			if (fp_is_small(&y->Z, 1)) {
				fp_to_be(&pubs[i * 2 * plen], &y->X, pl);
				fp_to_be(&pubs[i * 2 * plen + plen], &y->Y, pl);
			} else {
				prj.push_back(i);
			}
			ok[i] = 1;
		}
	});
	if (mixed.load()) return -1;
	std::vector<uint32_t> prj_idx;
	for (auto &part : prj_parts) prj_idx.insert(prj_idx.end(), part.begin(), part.end());
	if (!prj_idx.empty()) {
		std::vector<uint8_t> pb(prj_idx.size() * 3 * plen), abuf(prj_idx.size() * 2 * plen);
		std::vector<int8_t> st(prj_idx.size());
		parallel_for((uint32_t)prj_idx.size(), [&](uint32_t lo, uint32_t hi, unsigned) {
			for (uint32_t k = lo; k < hi; k++) {
				const eccb200_prj_pt *p = &pub_keys[prj_idx[k]]->y;
				fp_to_be(&pb[k * 3 * plen], &p->X, pl);
				fp_to_be(&pb[k * 3 * plen + plen], &p->Y, pl);
				fp_to_be(&pb[k * 3 * plen + 2 * plen], &p->Z, pl);
			}
		});
		if (eccb200_prj_pt_unique_batch(eng, (uint32_t)prj_idx.size(), pb.data(), abuf.data(), st.data())) return -1;
		for (size_t k = 0; k < prj_idx.size(); k++) {
			uint32_t i = prj_idx[k];
			if (st[k] == 0) memcpy(&pubs[i * 2 * plen], &abuf[k * 2 * plen], 2 * plen);
			else ok[i] = 0; /* key off the curve; a key at infinity gives W' = sG in the reference — rejected here,
					 * the same documented divergence as for ECFSDSA */
		}
	}
	for (uint32_t i = 0; i < num; i++)
		if (!ok[i]) { /* keep the batch launchable: rejected slots multiply the generator by zero */
			memset(&ab[i * 2 * qlen], 0, 2 * qlen);
			gen_to_be(&pubs[i * 2 * plen], ci);
		}
	if (eccb200_double_smul_batch(eng, num, ab, pubs, wout, status)) return -1; /* 4. W' = sG + eY (:495-498) */
	g_verifies += num;
	std::vector<int8_t> verdict(num, -1);
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
		for (uint32_t i = lo; i < hi; i++) {
			if (!ok[i] || status[i] != 0) continue; /* infinity: prj_pt_unique fails (:498) */
			uint8_t rp[128];
			const unsigned char *in[4] = { &wout[i * 2 * plen], optimized ? m[i] : &wout[i * 2 * plen + plen],
						       optimized ? nullptr : m[i], nullptr };
			uint32_t il[3] = { (uint32_t)plen, optimized ? m_len[i] : (uint32_t)plen, optimized ? 0u : m_len[i] };
			if (hm->hfunc_scattered(in, il, rp)) continue;          /* 5. r' = H(W'x [|| W'y] || m) (:500-520) */
			verdict[i] = memcmp(rp, s[i], hlen) == 0 ? 0 : -1;      /* 6. r == r' (sig/ecsdsa_common.c:606) */
		}
	});
	int all = 0;
	for (uint32_t i = 0; i < num; i++)
		if (verdict[i]) all = -1;
	t_verdicts.assign(verdict.begin(), verdict.end());
	return all;
}

/*
 * ECKCDSA (sig/eckcdsa.c:543-832): r_len = min(|H|, qlen); s in ]0, q[; h = H(z || m) with z = the first block_size
 * bytes of Y_x || Y_y || 0...; e = OS2I(r XOR rightmost(h)) mod q; W' = sY + eG; r' = rightmost(H(W'_x)) == r.  Same
 * split as ECSDSA: host for the hashes and the mod-q scalar, one device launch (eccb200_double_smul_batch with a = e
 * on G and b = s on Y) for the whole batch.
 */
static int verify_batch_eckcdsa(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
				const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
				const uint8_t **adata)
{
	t_verdicts.assign(num, -1);
	if (num == 0) return -1;
	if (!s || !s_len || !pub_keys || !m || !m_len) return -1;
	if (adata)
		for (uint32_t i = 0; i < num; i++)
			if (adata[i]) return -1;
	get_hash_fn get_hash = resolve_get_hash();
	if (!get_hash) return -1;
	const HashMappingHead *hm = nullptr;
	if (get_hash(hash_type, &hm) || !hm || !hm->hfunc_scattered) return -1;
	const uint32_t hlen = hm->digest_size, zlen = hm->block_size;
	if (hlen == 0 || hlen > 128 || zlen == 0) return -1;
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
	const size_t rlen = std::min<size_t>(hlen, qlen), siglen = rlen + qlen; /* ECKCDSA_R_LEN / _SIGLEN (sig/eckcdsa.h:28-31) */
	const size_t shift = hlen > rlen ? hlen - rlen : 0;
	uint8_t *ab = engine.slot->st[0].get(num * 2 * qlen), *pubs = engine.slot->st[1].get(num * 2 * plen),
		*wout = engine.slot->st[2].get(num * 2 * plen);
	int8_t *status = (int8_t *)engine.slot->st[3].get(num);
	if (!ab || !pubs || !wout || !status) return -1;
	uint8_t qbe[72];
	words_to_be(qbe, (int)qlen, ci->q);
	std::vector<uint8_t> ok(num, 0);
	std::atomic<int> mixed{ 0 };
	std::vector<std::vector<uint32_t>> prj_parts(64);
	/* pass 1: struct and range checks, affine keys (z needs them) */
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned t) {
		std::vector<uint32_t> &prj = prj_parts[t];
		for (uint32_t i = lo; i < hi; i++) {
			memset(&ab[i * 2 * qlen], 0, 2 * qlen);
			memset(&pubs[i * 2 * plen], 0, 2 * plen);
			const eccb200_ec_pub_key *pk = pub_keys[i];
			if (!pk || pk->magic != kPubKeyMagic || pk->key_type != sig_type || !pt_ok(&pk->y)) continue;
			if (!s[i] || (!m[i] && m_len[i])) continue;
			const CurveInfo *c = identify(&pk->y);
			if (!c) continue;
			if (c != ci) {
				mixed.store(1);
				continue;
			}
			if (s_len[i] != siglen) continue;                       /* 1. (:589) */
			const uint8_t *sb = s[i] + rlen;
			bool zero = true;
			for (size_t j = 0; j < qlen; j++) zero = zero && sb[j] == 0;
			if (zero || memcmp(sb, qbe, qlen) >= 0) continue;       /* 2. s in ]0, q[ (:592-595) */
			const eccb200_prj_pt *y = &pk->y;
			if (fp_is_small(&y->Z, 1)) {
				fp_to_be(&pubs[i * 2 * plen], &y->X, pl);
				fp_to_be(&pubs[i * 2 * plen + plen], &y->Y, pl);
			} else {
				prj.push_back(i);
			}
			ok[i] = 1;
		}
	});
	if (mixed.load()) return -1;
	std::vector<uint32_t> prj_idx;
	for (auto &part : prj_parts) prj_idx.insert(prj_idx.end(), part.begin(), part.end());
	if (!prj_idx.empty()) {
		std::vector<uint8_t> pb(prj_idx.size() * 3 * plen), abuf(prj_idx.size() * 2 * plen);
		std::vector<int8_t> st(prj_idx.size());
		parallel_for((uint32_t)prj_idx.size(), [&](uint32_t lo, uint32_t hi, unsigned) {
			for (uint32_t k = lo; k < hi; k++) {
				const eccb200_prj_pt *p = &pub_keys[prj_idx[k]]->y;
				fp_to_be(&pb[k * 3 * plen], &p->X, pl);
				fp_to_be(&pb[k * 3 * plen + plen], &p->Y, pl);
				fp_to_be(&pb[k * 3 * plen + 2 * plen], &p->Z, pl);
			}
		});
		if (eccb200_prj_pt_unique_batch(eng, (uint32_t)prj_idx.size(), pb.data(), abuf.data(), st.data())) return -1;
		for (size_t k = 0; k < prj_idx.size(); k++) {
			uint32_t i = prj_idx[k];
			if (st[k] == 0) memcpy(&pubs[i * 2 * plen], &abuf[k * 2 * plen], 2 * plen);
			else ok[i] = 0; /* off the curve, or infinity: prj_pt_to_aff fails on it (:614) */
		}
	}
	/* pass 2: h = H(z || m), e = OS2I(r XOR rightmost(h)) mod q */
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
		std::vector<uint8_t> z(zlen);
		for (uint32_t i = lo; i < hi; i++) {
			if (!ok[i]) {
				memset(&ab[i * 2 * qlen], 0, 2 * qlen);
				gen_to_be(&pubs[i * 2 * plen], ci); /* keep the batch launchable */
				continue;
			}
			std::fill(z.begin(), z.end(), 0);
			memcpy(z.data(), &pubs[i * 2 * plen], std::min<size_t>(zlen, 2 * plen)); /* 3. z (:601-625) */
			uint8_t h[128], x[128];
			const unsigned char *in[3] = { z.data(), m[i], nullptr };
			uint32_t il[2] = { zlen, m_len[i] };
			if (hm->hfunc_scattered(in, il, h)) {
				ok[i] = 0;
				continue;
			}
			for (size_t j = 0; j < rlen; j++) x[j] = (uint8_t)(h[shift + j] ^ s[i][j]); /* 4.-5. (:754-762) */
			eccb200_nn t;
			be_to_nn(&t, x, (uint32_t)rlen);
			scalar_mod_to_be(&ab[i * 2 * qlen], &t, ci);                  /* e on G */
			memcpy(&ab[i * 2 * qlen + qlen], s[i] + rlen, qlen);         /* s on Y */
		}
	});
	if (eccb200_double_smul_batch(eng, num, ab, pubs, wout, status)) return -1; /* 6. W' = sY + eG (:770-773) */
	g_verifies += num;
	std::vector<int8_t> verdict(num, -1);
	parallel_for(num, [&](uint32_t lo, uint32_t hi, unsigned) {
		for (uint32_t i = lo; i < hi; i++) {
			if (!ok[i] || status[i] != 0) continue; /* infinity: prj_pt_unique fails (:773) */
			uint8_t rp[128];
			const unsigned char *in[2] = { &wout[i * 2 * plen], nullptr };
			uint32_t il[1] = { (uint32_t)plen };
			if (hm->hfunc_scattered(in, il, rp)) continue;                /* 7. r' = H(W'_x) (:778-784) */
			verdict[i] = memcmp(rp + shift, s[i], rlen) == 0 ? 0 : -1;   /* 8.-9. (:794-800) */
		}
	});
	int all = 0;
	for (uint32_t i = 0; i < num; i++)
		if (verdict[i]) all = -1;
	t_verdicts.assign(verdict.begin(), verdict.end());
	return all;
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
	(void)adata_len;
	if (sig_type != 3 /* ECSDSA */ && sig_type != 4 /* ECOSDSA */) return -1;
	return verify_batch_ecsdsa(sig_type == 4, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
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
	(void)adata_len;
	if (sig_type != 2 /* ECKCDSA */) return -1;
	return verify_batch_eckcdsa(s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
}

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
	const bool post_hash = (sig_type == 2 || sig_type == 3 || sig_type == 4); /* ECKCDSA / ECSDSA / ECOSDSA: these hash
										    * the recomputed point */
	bool ours = (post_hash || scheme_of(sig_type, &sc)) && !adata && adata_len == 0 && sig && pub_key &&
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
	if (sig_type == 2) return verify_batch_eckcdsa(sp, sl, pk, mp, ml, 1, sig_type, hash_type, nullptr);
	if (post_hash) return verify_batch_ecsdsa(sig_type == 4, sp, sl, pk, mp, ml, 1, sig_type, hash_type, nullptr);
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
	return scheme_of(sig_type, &sc) || sig_type == 2 || sig_type == 3 || sig_type == 4;
}

extern "C" int eccb200_dropin_ec_verify_batch(const uint8_t **s, const uint8_t *s_len,
					      const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
					      const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
					      const uint8_t **adata, const uint16_t *adata_len, void *scratch_pad_area,
					      uint32_t *scratch_pad_area_len)
{
	bool ours = batch_scheme_served(sig_type) && num > 0 && s && s_len && pub_keys && m && m_len &&
		    resolve_get_hash() != nullptr;
	if (ours && adata)
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
	if (sig_type == 2) return verify_batch_eckcdsa(s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
	if (sig_type == 3 || sig_type == 4)
		return verify_batch_ecsdsa(sig_type == 4, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
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
