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
#include <cstdlib>
#include <cstring>
#include <mutex>
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

std::mutex g_mu;
unsigned long long g_calls = 0; /* scalar multiplications served (eccb200_dropin_call_count) */
eccb200_ctx *g_ctx[32] = { nullptr }; /* indexed by the reference's ec_curve_type (< 32 here) */
int g_device = -1;

eccb200_ctx *engine_for(int curve_id)
{
	/* caller holds g_mu */
	if (g_device < 0) {
		const char *e = getenv("ECCB200_DEVICE");
		g_device = e ? atoi(e) : 0;
	}
	if (!g_ctx[curve_id]) {
		const char *w = getenv("ECCB200_COMB_WINDOW");
		if (eccb200_ctx_create(&g_ctx[curve_id], curve_id, g_device, w ? atoi(w) : 0)) return nullptr;
	}
	return g_ctx[curve_id];
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
	std::lock_guard<std::mutex> lk(g_mu);
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
	eccb200_ctx *eng = engine_for(ci->id);
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
	std::lock_guard<std::mutex> lk(g_mu);
	for (auto &c : g_ctx) {
		if (c) eccb200_ctx_destroy(c);
		c = nullptr;
	}
	g_device = device;
	return 0;
}

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

/* prj_pt_mul_blind (curves/prj_pt.c:1782-1822) adds a random multiple of the order to the scalar and calls
 * prj_pt_mul: the result is the same point, so the drop-in forwards to the same path. */
extern "C" int prj_pt_mul_blind(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in)
{
	return mul_batch(out, m, in, 1, nullptr);
}

extern "C" unsigned long long eccb200_dropin_call_count(void)
{
	std::lock_guard<std::mutex> lk(g_mu);
	return g_calls;
}

extern "C" uint32_t eccb200_dropin_last_verdicts(int8_t *verdicts, uint32_t cap)
{
	uint32_t k = (uint32_t)t_verdicts.size();
	if (k > cap) k = cap;
	if (verdicts && k) memcpy(verdicts, t_verdicts.data(), k);
	return k;
}

/* fs = false: ECDSA / DECDSA (signature r || s, digest H(m));  fs = true: ECFSDSA (signature W_x || W_y || s, digest
 * H(W_x || W_y || m), sig/ecfsdsa.c:482,529) */
static int verify_batch_common(bool fs, const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
			       const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
			       const uint8_t **adata)
{
	t_verdicts.assign(num, -1);
	if (num == 0) return -1; /* the reference's implementations reject an empty batch (sig/ecfsdsa.c:740) */
	if (!s || !s_len || !pub_keys || !m || !m_len) return -1;
	if (fs ? (sig_type != 5 /* ECFSDSA */) : (sig_type != 1 /* ECDSA */ && sig_type != 14 /* DECDSA */)) return -1;
	if (adata) /* ECDSA takes no ancillary data: every entry must be NULL (sig/ecdsa.c:76-83) */
		for (uint32_t i = 0; i < num; i++)
			if (adata[i]) return -1;
	/* the reference's own hash (src/hash stays host-side): resolve get_hash_by_type from the application */
	typedef int (*get_hash_fn)(int, const HashMappingHead **);
	static get_hash_fn get_hash = (get_hash_fn)dlsym(RTLD_DEFAULT, "get_hash_by_type");
	if (!get_hash) return -1;
	const HashMappingHead *hm = nullptr;
	if (get_hash(hash_type, &hm) || !hm || !hm->hfunc_scattered) return -1;
	const uint32_t hlen = hm->digest_size;

	std::lock_guard<std::mutex> lk(g_mu);
	const CurveInfo *ci = nullptr;
	std::vector<uint8_t> ok(num, 0);
	for (uint32_t i = 0; i < num; i++) {
		const eccb200_ec_pub_key *pk = pub_keys[i];
		if (!pk || pk->magic != kPubKeyMagic || pk->key_type != sig_type || !pt_ok(&pk->y)) continue;
		if (!s[i] || (!m[i] && m_len[i])) continue;
		const CurveInfo *c = identify(&pk->y);
		if (!c) continue;
		if (ci && c != ci) return -1; /* all keys must share the curve parameters (sig/ecfsdsa.c:711) */
		ci = c;
		ok[i] = 1;
	}
	if (!ci) return -1;
	eccb200_ctx *eng = engine_for(ci->id);
	if (!eng) return -1;
	const int pl = ci->plen;
	const size_t plen = (size_t)ci->plen, qlen = (size_t)ci->qlen;
	const size_t siglen = fs ? 2 * plen + qlen : 2 * qlen;
	std::vector<uint8_t> sigs(num * siglen, 0), pubs(num * 2 * plen, 0), dig(num * (size_t)hlen, 0);
	/* public keys whose y is not already (x, y, 1) go through the batched prj_pt_unique */
	std::vector<uint32_t> prj_idx;
	for (uint32_t i = 0; i < num; i++) {
		if (!ok[i]) continue;
		if (s_len[i] != siglen) { /* siglen check, sig/ecdsa_common.c:645, sig/ecfsdsa.c:447 */
			ok[i] = 0;
			continue;
		}
		memcpy(&sigs[i * siglen], s[i], siglen);
		const unsigned char *inputs[3] = { fs ? s[i] : m[i], fs ? m[i] : nullptr, nullptr };
		uint32_t ilens[2] = { fs ? (uint32_t)(2 * plen) : m_len[i], fs ? m_len[i] : 0 };
		if (hm->hfunc_scattered(inputs, ilens, &dig[i * (size_t)hlen])) {
			ok[i] = 0;
			continue;
		}
		const eccb200_prj_pt *y = &pub_keys[i]->y;
		if (fp_is_small(&y->Z, 1)) {
			fp_to_be(&pubs[i * 2 * plen], &y->X, pl);
			fp_to_be(&pubs[i * 2 * plen + plen], &y->Y, pl);
		} else {
			prj_idx.push_back(i);
		}
	}
	if (!prj_idx.empty()) {
		std::vector<uint8_t> pb(prj_idx.size() * 3 * plen), ab(prj_idx.size() * 2 * plen);
		std::vector<int8_t> st(prj_idx.size());
		for (size_t k = 0; k < prj_idx.size(); k++) {
			const eccb200_prj_pt *p = &pub_keys[prj_idx[k]]->y;
			fp_to_be(&pb[k * 3 * plen], &p->X, pl);
			fp_to_be(&pb[k * 3 * plen + plen], &p->Y, pl);
			fp_to_be(&pb[k * 3 * plen + 2 * plen], &p->Z, pl);
		}
		if (eccb200_prj_pt_unique_batch(eng, (uint32_t)prj_idx.size(), pb.data(), ab.data(), st.data())) return -1;
		for (size_t k = 0; k < prj_idx.size(); k++) {
			uint32_t i = prj_idx[k];
			if (st[k] != 0) ok[i] = 0;
			else memcpy(&pubs[i * 2 * plen], &ab[k * 2 * plen], 2 * plen);
		}
	}
	std::vector<int8_t> verdict(num, -1);
	if (fs ? eccb200_ecfsdsa_verify_batch(eng, num, sigs.data(), pubs.data(), dig.data(), hlen, verdict.data())
	       : eccb200_ecdsa_verify_batch(eng, num, sigs.data(), pubs.data(), dig.data(), hlen, verdict.data()))
		return -1;
	int all = 0;
	for (uint32_t i = 0; i < num; i++) {
		if (!ok[i]) verdict[i] = -1;
		if (verdict[i]) all = -1;
	}
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
	return verify_batch_common(false, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
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
	return verify_batch_common(true, s, s_len, pub_keys, m, m_len, num, sig_type, hash_type, adata);
}
