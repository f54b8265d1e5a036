/*
 * oracle/ref_shim.c — flat batch wrappers around the UNMODIFIED reference (libecc) API.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with the reference's own sources (where they lie
 * under /root/reference, see oracle/Makefile) into oracle/_ref/libecc_ref.so.  It contains no arithmetic of its
 * own: every result comes from the reference's prj_pt_mul / prj_pt_unique / ec_verify / ec_sign / hash code.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs load it, and only as the checker / CPU baseline.
 *
 * Reference entry points used (paths relative to /root/reference/src):
 *   ec_get_curve_params_by_name  curves/curves.c:25      import_params           curves/ec_params.c:24
 *   prj_pt_import_from_aff_buf   curves/prj_pt.c:511     prj_pt_mul              curves/prj_pt.c:1759
 *   prj_pt_iszero                curves/prj_pt.c:107     prj_pt_export_to_aff_buf curves/prj_pt.c:600
 *   ec_pub_key_import_from_aff_buf sig/ec_key.c          ec_verify               sig/sig_algs.c:655
 *   ec_key_pair_import_from_priv_key_buf sig/ec_key.c    ec_sign                 sig/sig_algs.c:497
 *   get_hash_by_name             hash/hash_algs.c
 */
#include "libsig.h"
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>

typedef struct {
	ec_params params;
	u32 plen, qlen;
} ref_curve;

static int ref_load_curve(ref_curve *c, const char *name)
{
	const ec_str_params *sp = NULL;
	size_t l = strlen(name);
	if (l > 250) return -1;
	if (ec_get_curve_params_by_name((const u8 *)name, (u8)(l + 1), &sp) || sp == NULL) return -1;
	if (import_params(&c->params, sp)) return -1;
	c->plen = (u32)BYTECEIL(c->params.ec_fp.p_bitlen);
	c->qlen = (u32)BYTECEIL(c->params.ec_gen_order_bitlen);
	return 0;
}

/* Curve constants as big-endian byte strings (each buffer must hold plen or qlen bytes). */
int ref_curve_info(const char *name, uint32_t *plen, uint32_t *qlen, uint8_t *p, uint8_t *q, uint8_t *a,
		   uint8_t *b, uint8_t *gx, uint8_t *gy)
{
	ref_curve c;
	u8 buf[2 * 128];
	if (ref_load_curve(&c, name)) return -1;
	*plen = c.plen;
	*qlen = c.qlen;
	if (nn_export_to_buf(p, (u16)c.plen, &c.params.ec_fp.p)) return -1;
	if (nn_export_to_buf(q, (u16)c.qlen, &c.params.ec_gen_order)) return -1;
	if (fp_export_to_buf(a, (u16)c.plen, &c.params.ec_curve.a)) return -1;
	if (fp_export_to_buf(b, (u16)c.plen, &c.params.ec_curve.b)) return -1;
	if (prj_pt_export_to_aff_buf(&c.params.ec_gen, buf, 2 * c.plen)) return -1;
	memcpy(gx, buf, c.plen);
	memcpy(gy, buf + c.plen, c.plen);
	return 0;
}

/* ---------------------------------------------------------------- scalar multiplication */
typedef struct {
	const ref_curve *c;
	uint32_t lo, hi;
	const uint8_t *scalars;
	uint32_t slen;
	const uint8_t *points;
	uint8_t *out;
	int8_t *status;
} smul_job;

static void *smul_worker(void *arg)
{
	smul_job *j = (smul_job *)arg;
	const ref_curve *c = j->c;
	u32 plen = c->plen;
	for (uint32_t i = j->lo; i < j->hi; i++) {
		nn k;
		prj_pt in, out;
		int iszero = 0;
		uint8_t *o = j->out + (size_t)i * 2 * plen;
		memset(o, 0, 2 * plen);
		j->status[i] = -1;
		if (nn_init_from_buf(&k, j->scalars + (size_t)i * j->slen, (u16)j->slen)) continue;
		if (j->points) {
			/* import checks that the point is on the curve (curves/prj_pt.c:541-545) */
			if (prj_pt_import_from_aff_buf(&in, j->points + (size_t)i * 2 * plen, (u16)(2 * plen),
						       &c->params.ec_curve))
				continue;
		} else {
			if (prj_pt_copy(&in, &c->params.ec_gen)) continue;
		}
		if (prj_pt_mul(&out, &k, &in)) continue;
		if (prj_pt_iszero(&out, &iszero)) continue;
		if (iszero) {
			j->status[i] = 1;
			continue;
		}
		/* prj_pt_export_to_aff_buf normalises through prj_pt_to_aff (fp_inv) */
		if (prj_pt_export_to_aff_buf(&out, o, 2 * plen)) continue;
		j->status[i] = 0;
	}
	return NULL;
}

/*
 * out[i] = affine big-endian (x || y) of scalars[i] * (points ? points[i] : G)
 * status[i]: 0 finite result, 1 point at infinity (out zeroed), -1 reference returned an error.
 */
int ref_prj_pt_mul_batch(const char *curve, uint32_t n, const uint8_t *scalars, uint32_t slen,
			 const uint8_t *points, uint8_t *out, int8_t *status, int nthreads)
{
	ref_curve c;
	if (ref_load_curve(&c, curve)) return -1;
	if (nthreads < 1) nthreads = 1;
	if ((uint32_t)nthreads > n) nthreads = n ? (int)n : 1;
	pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
	smul_job *jobs = (smul_job *)calloc((size_t)nthreads, sizeof(smul_job));
	for (int t = 0; t < nthreads; t++) {
		jobs[t].c = &c;
		jobs[t].lo = (uint32_t)(((uint64_t)n * (uint64_t)t) / (uint64_t)nthreads);
		jobs[t].hi = (uint32_t)(((uint64_t)n * (uint64_t)(t + 1)) / (uint64_t)nthreads);
		jobs[t].scalars = scalars;
		jobs[t].slen = slen;
		jobs[t].points = points;
		jobs[t].out = out;
		jobs[t].status = status;
		pthread_create(&th[t], NULL, smul_worker, &jobs[t]);
	}
	for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	free(th);
	free(jobs);
	return 0;
}

/* ---------------------------------------------------------------- ECDSA verify / sign */
static int ref_hash_type(const char *hash, hash_alg_type *t, u8 *dlen)
{
	const hash_mapping *hm = NULL;
	if (get_hash_by_name(hash, &hm) || hm == NULL) return -1;
	*t = hm->type;
	*dlen = hm->digest_size;
	return 0;
}

int ref_hash(const char *hash, const uint8_t *msg, uint32_t len, uint8_t *out, uint32_t *outlen)
{
	const hash_mapping *hm = NULL;
	const u8 *in[2] = { msg, NULL };
	u32 ilens[1] = { len };
	if (get_hash_by_name(hash, &hm) || hm == NULL) return -1;
	if (hm->hfunc_scattered(in, ilens, out)) return -1;
	*outlen = hm->digest_size;
	return 0;
}

typedef struct {
	const ref_curve *c;
	hash_alg_type ht;
	uint32_t lo, hi;
	const uint8_t *sigs;
	const uint8_t *pubkeys;
	const uint8_t *privkeys;
	const uint8_t *msgs;
	const uint64_t *off;
	uint8_t *sigs_out;
	uint8_t *pubkeys_out;
	int8_t *verdict;
	ec_alg_type alg;   /* ECDSA or ECFSDSA */
	uint32_t siglen;   /* 2*qlen (ECDSA) or 2*plen + qlen (ECFSDSA, sig/ecfsdsa.h) */
} sig_job;

static void *verify_worker(void *arg)
{
	sig_job *j = (sig_job *)arg;
	const ref_curve *c = j->c;
	for (uint32_t i = j->lo; i < j->hi; i++) {
		ec_pub_key pk;
		j->verdict[i] = -1;
		if (ec_pub_key_import_from_aff_buf(&pk, &c->params, j->pubkeys + (size_t)i * 2 * c->plen,
						   (u8)(2 * c->plen), j->alg))
			continue;
		if (ec_verify(j->sigs + (size_t)i * j->siglen, (u8)j->siglen, &pk, j->msgs + j->off[i],
			      (u32)(j->off[i + 1] - j->off[i]), j->alg, j->ht, NULL, 0))
			continue;
		j->verdict[i] = 0;
	}
	return NULL;
}

static void *sign_worker(void *arg)
{
	sig_job *j = (sig_job *)arg;
	const ref_curve *c = j->c;
	for (uint32_t i = j->lo; i < j->hi; i++) {
		ec_key_pair kp;
		j->verdict[i] = -1;
		if (ec_key_pair_import_from_priv_key_buf(&kp, &c->params, j->privkeys + (size_t)i * c->qlen,
							 (u8)c->qlen, j->alg))
			continue;
		if (ec_pub_key_export_to_aff_buf(&kp.pub_key, j->pubkeys_out + (size_t)i * 2 * c->plen,
						 (u8)(2 * c->plen)))
			continue;
		if (ec_sign(j->sigs_out + (size_t)i * j->siglen, (u8)j->siglen, &kp, j->msgs + j->off[i],
			    (u32)(j->off[i + 1] - j->off[i]), j->alg, j->ht, NULL, 0))
			continue;
		j->verdict[i] = 0;
	}
	return NULL;
}

static int run_sig_jobs(void *(*fn)(void *), sig_job *proto, uint32_t n, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	if ((uint32_t)nthreads > n) nthreads = n ? (int)n : 1;
	pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
	sig_job *jobs = (sig_job *)calloc((size_t)nthreads, sizeof(sig_job));
	for (int t = 0; t < nthreads; t++) {
		jobs[t] = *proto;
		jobs[t].lo = (uint32_t)(((uint64_t)n * (uint64_t)t) / (uint64_t)nthreads);
		jobs[t].hi = (uint32_t)(((uint64_t)n * (uint64_t)(t + 1)) / (uint64_t)nthreads);
		pthread_create(&th[t], NULL, fn, &jobs[t]);
	}
	for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	free(th);
	free(jobs);
	return 0;
}

/* verdict[i] = 0 iff the reference's ec_verify(…, ECDSA, hash, NULL, 0) returns 0, else -1 (also when the
 * public key fails to import, i.e. is not on the curve). Message i is msgs[off[i] .. off[i+1]). */
int ref_ecdsa_verify_batch(const char *curve, const char *hash, uint32_t n, const uint8_t *sigs,
			   const uint8_t *pubkeys, const uint8_t *msgs, const uint64_t *off, int8_t *verdict,
			   int nthreads)
{
	ref_curve c;
	sig_job p;
	u8 dlen;
	memset(&p, 0, sizeof(p));
	if (ref_load_curve(&c, curve)) return -1;
	if (ref_hash_type(hash, &p.ht, &dlen)) return -1;
	p.c = &c;
	p.sigs = sigs;
	p.pubkeys = pubkeys;
	p.msgs = msgs;
	p.off = off;
	p.verdict = verdict;
	p.alg = ECDSA;
	p.siglen = 2 * c.qlen;
	return run_sig_jobs(verify_worker, &p, n, nthreads);
}

/* Sign message i with private key i (qlen big-endian bytes, must be in [1,q-1]) using the reference's ec_sign
 * (fresh random nonce from /dev/urandom); also exports the matching affine public key. status[i] 0 / -1. */
int ref_ecdsa_sign_batch(const char *curve, const char *hash, uint32_t n, const uint8_t *privkeys,
			 const uint8_t *msgs, const uint64_t *off, uint8_t *sigs_out, uint8_t *pubkeys_out,
			 int8_t *status, int nthreads)
{
	ref_curve c;
	sig_job p;
	u8 dlen;
	memset(&p, 0, sizeof(p));
	if (ref_load_curve(&c, curve)) return -1;
	if (ref_hash_type(hash, &p.ht, &dlen)) return -1;
	p.c = &c;
	p.privkeys = privkeys;
	p.msgs = msgs;
	p.off = off;
	p.sigs_out = sigs_out;
	p.pubkeys_out = pubkeys_out;
	p.verdict = status;
	p.alg = ECDSA;
	p.siglen = 2 * c.qlen;
	return run_sig_jobs(sign_worker, &p, n, nthreads);
}

/* The same two for ECFSDSA (sig/ecfsdsa.c): signatures are r || s with r = W_x || W_y (2*plen bytes) and s (qlen). */
int ref_ecfsdsa_verify_batch(const char *curve, const char *hash, uint32_t n, const uint8_t *sigs,
			     const uint8_t *pubkeys, const uint8_t *msgs, const uint64_t *off, int8_t *verdict,
			     int nthreads)
{
	ref_curve c;
	sig_job p;
	u8 dlen;
	memset(&p, 0, sizeof(p));
	if (ref_load_curve(&c, curve)) return -1;
	if (ref_hash_type(hash, &p.ht, &dlen)) return -1;
	p.c = &c;
	p.sigs = sigs;
	p.pubkeys = pubkeys;
	p.msgs = msgs;
	p.off = off;
	p.verdict = verdict;
	p.alg = ECFSDSA;
	p.siglen = 2 * c.plen + c.qlen;
	return run_sig_jobs(verify_worker, &p, n, nthreads);
}

int ref_ecfsdsa_sign_batch(const char *curve, const char *hash, uint32_t n, const uint8_t *privkeys,
			   const uint8_t *msgs, const uint64_t *off, uint8_t *sigs_out, uint8_t *pubkeys_out,
			   int8_t *status, int nthreads)
{
	ref_curve c;
	sig_job p;
	u8 dlen;
	memset(&p, 0, sizeof(p));
	if (ref_load_curve(&c, curve)) return -1;
	if (ref_hash_type(hash, &p.ht, &dlen)) return -1;
	p.c = &c;
	p.privkeys = privkeys;
	p.msgs = msgs;
	p.off = off;
	p.sigs_out = sigs_out;
	p.pubkeys_out = pubkeys_out;
	p.verdict = status;
	p.alg = ECFSDSA;
	p.siglen = 2 * c.plen + c.qlen;
	return run_sig_jobs(sign_worker, &p, n, nthreads);
}

/* The same two for BIP0340 (sig/bip0340.c): signatures are r || s with r = x(kG) (plen bytes) and s (qlen).  The
 * reference's signer draws its auxiliary randomness from /dev/urandom and negates the key / nonce as BIP0340 says. */
int ref_bip0340_verify_batch(const char *curve, const char *hash, uint32_t n, const uint8_t *sigs,
			     const uint8_t *pubkeys, const uint8_t *msgs, const uint64_t *off, int8_t *verdict,
			     int nthreads)
{
	ref_curve c;
	sig_job p;
	u8 dlen;
	memset(&p, 0, sizeof(p));
	if (ref_load_curve(&c, curve)) return -1;
	if (ref_hash_type(hash, &p.ht, &dlen)) return -1;
	p.c = &c;
	p.sigs = sigs;
	p.pubkeys = pubkeys;
	p.msgs = msgs;
	p.off = off;
	p.verdict = verdict;
	p.alg = BIP0340;
	p.siglen = c.plen + c.qlen;
	return run_sig_jobs(verify_worker, &p, n, nthreads);
}

int ref_bip0340_sign_batch(const char *curve, const char *hash, uint32_t n, const uint8_t *privkeys,
			   const uint8_t *msgs, const uint64_t *off, uint8_t *sigs_out, uint8_t *pubkeys_out,
			   int8_t *status, int nthreads)
{
	ref_curve c;
	sig_job p;
	u8 dlen;
	memset(&p, 0, sizeof(p));
	if (ref_load_curve(&c, curve)) return -1;
	if (ref_hash_type(hash, &p.ht, &dlen)) return -1;
	p.c = &c;
	p.privkeys = privkeys;
	p.msgs = msgs;
	p.off = off;
	p.sigs_out = sigs_out;
	p.pubkeys_out = pubkeys_out;
	p.verdict = status;
	p.alg = BIP0340;
	p.siglen = c.plen + c.qlen;
	return run_sig_jobs(sign_worker, &p, n, nthreads);
}

/* Generic forms for any scheme of the reference, named like its ec_alg_type ("ECSDSA", "ECOSDSA", "ECKCDSA", ...):
 * ec_sign / ec_verify per item; the signature length comes from ec_get_sig_len and is returned through *siglen_out. */
static int ref_alg_by_name(const char *name, ec_alg_type *alg)
{
	static const struct { const char *n; ec_alg_type a; } tab[] = {
		{ "ECDSA", ECDSA }, { "ECKCDSA", ECKCDSA }, { "ECSDSA", ECSDSA }, { "ECOSDSA", ECOSDSA },
		{ "ECFSDSA", ECFSDSA }, { "ECGDSA", ECGDSA }, { "ECRDSA", ECRDSA }, { "DECDSA", DECDSA },
		{ "BIP0340", BIP0340 },
	};
	for (unsigned i = 0; i < sizeof(tab) / sizeof(tab[0]); i++)
		if (!strcmp(name, tab[i].n)) {
			*alg = tab[i].a;
			return 0;
		}
	return -1;
}

int ref_sig_len(const char *curve, const char *alg_name, const char *hash, uint32_t *siglen_out)
{
	ref_curve c;
	ec_alg_type alg;
	hash_alg_type ht;
	u8 dlen, sl = 0;
	if (ref_load_curve(&c, curve) || ref_alg_by_name(alg_name, &alg) || ref_hash_type(hash, &ht, &dlen)) return -1;
	if (ec_get_sig_len(&c.params, alg, ht, &sl)) return -1;
	*siglen_out = sl;
	return 0;
}

int ref_sig_verify_batch(const char *curve, const char *alg_name, const char *hash, uint32_t n, const uint8_t *sigs,
			 const uint8_t *pubkeys, const uint8_t *msgs, const uint64_t *off, int8_t *verdict, int nthreads)
{
	ref_curve c;
	sig_job p;
	u8 dlen, sl = 0;
	memset(&p, 0, sizeof(p));
	if (ref_load_curve(&c, curve) || ref_alg_by_name(alg_name, &p.alg) || ref_hash_type(hash, &p.ht, &dlen)) return -1;
	if (ec_get_sig_len(&c.params, p.alg, p.ht, &sl)) return -1;
	p.c = &c;
	p.sigs = sigs;
	p.pubkeys = pubkeys;
	p.msgs = msgs;
	p.off = off;
	p.verdict = verdict;
	p.siglen = sl;
	return run_sig_jobs(verify_worker, &p, n, nthreads);
}

int ref_sig_sign_batch(const char *curve, const char *alg_name, const char *hash, uint32_t n, const uint8_t *privkeys,
		       const uint8_t *msgs, const uint64_t *off, uint8_t *sigs_out, uint8_t *pubkeys_out, int8_t *status,
		       int nthreads)
{
	ref_curve c;
	sig_job p;
	u8 dlen, sl = 0;
	memset(&p, 0, sizeof(p));
	if (ref_load_curve(&c, curve) || ref_alg_by_name(alg_name, &p.alg) || ref_hash_type(hash, &p.ht, &dlen)) return -1;
	if (ec_get_sig_len(&c.params, p.alg, p.ht, &sl)) return -1;
	p.c = &c;
	p.privkeys = privkeys;
	p.msgs = msgs;
	p.off = off;
	p.sigs_out = sigs_out;
	p.pubkeys_out = pubkeys_out;
	p.verdict = status;
	p.siglen = sl;
	return run_sig_jobs(sign_worker, &p, n, nthreads);
}

/* The reference's own batch entry point, ec_verify_batch(…, ECFSDSA, …) (sig/sig_algs.c:675 -> sig/ecfsdsa.c:1057):
 * one 0 / -1 answer for the whole batch.  use_scratch = 0: no scratch pad, the reference verifies the signatures one
 * after the other (sig/ecfsdsa.c:711); use_scratch = 1: its Bos-Coster multi-scalar multiplication (:842). */
static int verify_batch_all(ec_alg_type alg, const char *curve, const char *hash, uint32_t n, const uint8_t *sigs,
			    const uint8_t *pubkeys, const uint8_t *msgs, const uint64_t *off, int use_scratch)
{
	ref_curve c;
	hash_alg_type ht;
	u8 dlen;
	int ret = -1;
	if (ref_load_curve(&c, curve) || ref_hash_type(hash, &ht, &dlen) || n == 0) return -2;
	const uint32_t siglen = (alg == BIP0340 ? c.plen : 2 * c.plen) + c.qlen;
	ec_pub_key *pks = (ec_pub_key *)calloc(n, sizeof(ec_pub_key));
	const ec_pub_key **pkp = (const ec_pub_key **)calloc(n, sizeof(void *));
	const u8 **sp = (const u8 **)calloc(n, sizeof(void *)), **mp = (const u8 **)calloc(n, sizeof(void *));
	u8 *sl = (u8 *)calloc(n, 1);
	u32 *ml = (u32 *)calloc(n, sizeof(u32));
	u32 scratch_len = 0;
	verify_batch_scratch_pad *scratch = NULL;
	for (uint32_t i = 0; i < n; i++) {
		if (ec_pub_key_import_from_aff_buf(&pks[i], &c.params, pubkeys + (size_t)i * 2 * c.plen, (u8)(2 * c.plen), alg))
			goto out;
		pkp[i] = &pks[i];
		sp[i] = sigs + (size_t)i * siglen;
		sl[i] = (u8)siglen;
		mp[i] = msgs + off[i];
		ml[i] = (u32)(off[i + 1] - off[i]);
	}
	/* with a scratch pad the reference runs its Bos-Coster batch algorithm; it wants (2n + 1) entries
	 * (sig/ecfsdsa.c:898) */
	if (use_scratch) {
		scratch_len = (u32)((2 * (size_t)n + 1) * sizeof(verify_batch_scratch_pad));
		scratch = (verify_batch_scratch_pad *)calloc(1, scratch_len);
	}
	ret = ec_verify_batch(sp, sl, pkp, mp, ml, n, alg, ht, NULL, NULL, scratch, &scratch_len) ? -1 : 0;
out:
	free(pks); free(pkp); free(sp); free(mp); free(sl); free(ml); free(scratch);
	return ret;
}

int ref_ecfsdsa_verify_batch_all(const char *curve, const char *hash, uint32_t n, const uint8_t *sigs,
				 const uint8_t *pubkeys, const uint8_t *msgs, const uint64_t *off, int use_scratch)
{
	return verify_batch_all(ECFSDSA, curve, hash, n, sigs, pubkeys, msgs, off, use_scratch);
}

/* the same through the BIP0340 entry of ec_sig_maps[] (_bip0340_verify_batch, sig/bip0340.c:1040) */
int ref_bip0340_verify_batch_all(const char *curve, const char *hash, uint32_t n, const uint8_t *sigs,
				 const uint8_t *pubkeys, const uint8_t *msgs, const uint64_t *off, int use_scratch)
{
	return verify_batch_all(BIP0340, curve, hash, n, sigs, pubkeys, msgs, off, use_scratch);
}

/* ---------------------------------------------------------------- structured key / signature records (sig/ec_key.c) */

/* ec_key_pair_import_from_priv_key_buf + the two structured exports.  Returns the import's code; lengths out. */
int ref_structured_keygen(const char *curve, const uint8_t *priv, uint32_t priv_len, uint8_t *priv_rec,
			  uint32_t *priv_rec_len, uint8_t *pub_rec, uint32_t *pub_rec_len)
{
	ref_curve c;
	ec_key_pair kp;
	if (ref_load_curve(&c, curve)) return -2;
	if (ec_key_pair_import_from_priv_key_buf(&kp, &c.params, priv, (u8)priv_len, ECDSA)) return -1;
	*priv_rec_len = EC_STRUCTURED_PRIV_KEY_EXPORT_SIZE(&kp.priv_key);
	*pub_rec_len = EC_STRUCTURED_PUB_KEY_EXPORT_SIZE(&kp.pub_key);
	if (ec_structured_priv_key_export_to_buf(&kp.priv_key, priv_rec, (u8)*priv_rec_len)) return -3;
	if (ec_structured_pub_key_export_to_buf(&kp.pub_key, pub_rec, (u8)*pub_rec_len)) return -3;
	return 0;
}

/* ec_structured_pub_key_import_from_buf; on success the affine point (or is_inf = 1) */
int ref_structured_pub_import(const char *curve, const uint8_t *rec, uint32_t len, uint8_t *aff, int *is_inf)
{
	ref_curve c;
	ec_pub_key pk;
	int z = 0;
	if (ref_load_curve(&c, curve)) return -2;
	*is_inf = 0;
	memset(aff, 0, 2 * c.plen);
	if (ec_structured_pub_key_import_from_buf(&pk, &c.params, rec, (u8)len, ECDSA)) return -1;
	if (prj_pt_iszero(&pk.y, &z)) return -3;
	if (z) {
		*is_inf = 1;
		return 0;
	}
	if (prj_pt_export_to_aff_buf(&pk.y, aff, 2 * c.plen)) return -3;
	return 0;
}

/* What the reference's verify front end does with structured inputs (tests/ec_utils.c, verify_bin_file): import the
 * key record, import the signature record and require its (alg, hash, curve) to match, then ec_verify. */
int ref_structured_verify(const char *curve, const char *hash, const uint8_t *sig_rec, uint32_t sig_rec_len,
			  const uint8_t *pub_rec, uint32_t pub_rec_len, const uint8_t *msg, uint32_t mlen)
{
	ref_curve c;
	ec_pub_key pk;
	const hash_mapping *hm = NULL;
	u8 sig[EC_MAX_SIGLEN], name[MAX_CURVE_NAME_LEN];
	ec_alg_type st;
	hash_alg_type ht;
	u32 siglen;
	int same = 0;
	if (ref_load_curve(&c, curve)) return -2;
	if (get_hash_by_name(hash, &hm) || !hm) return -2;
	if (ec_structured_pub_key_import_from_buf(&pk, &c.params, pub_rec, (u8)pub_rec_len, ECDSA)) return -1;
	if (sig_rec_len < 3 || sig_rec_len - 3 > EC_MAX_SIGLEN) return -1;
	siglen = sig_rec_len - 3;
	if (ec_structured_sig_import_from_buf(sig, siglen, sig_rec, sig_rec_len, &st, &ht, name)) return -1;
	if (st != ECDSA || ht != hm->type) return -1;
	if (are_str_equal((const char *)name, (const char *)c.params.curve_name, &same) || !same) return -1;
	return ec_verify(sig, (u8)siglen, &pk, msg, mlen, ECDSA, hm->type, NULL, 0) ? -1 : 0;
}

/* sizeof/offsetof facts about the reference's structs (used to pin include/libecc_b200_dropin.h's mirror). */
int ref_abi_facts(uint64_t *out, uint32_t nmax)
{
	uint64_t f[] = {
		sizeof(nn), offsetof(nn, magic), offsetof(nn, wlen),
		sizeof(fp), offsetof(fp, ctx), offsetof(fp, magic),
		sizeof(prj_pt), offsetof(prj_pt, Y), offsetof(prj_pt, Z), offsetof(prj_pt, crv), offsetof(prj_pt, magic),
		sizeof(fp_ctx), offsetof(fp_ctx, p_bitlen), offsetof(fp_ctx, mpinv), offsetof(fp_ctx, r),
		offsetof(fp_ctx, r_square), offsetof(fp_ctx, magic),
		sizeof(ec_shortw_crv), offsetof(ec_shortw_crv, b), offsetof(ec_shortw_crv, a_monty),
		offsetof(ec_shortw_crv, order), offsetof(ec_shortw_crv, magic),
		sizeof(ec_params), offsetof(ec_params, ec_curve), offsetof(ec_params, ec_gen),
		offsetof(ec_params, ec_gen_order), offsetof(ec_params, ec_gen_order_bitlen),
		offsetof(ec_params, curve_name), offsetof(ec_params, curve_type),
		sizeof(ec_pub_key), offsetof(ec_pub_key, params), offsetof(ec_pub_key, y), offsetof(ec_pub_key, magic),
		NN_MAX_WORD_LEN,
	};
	uint32_t k = (uint32_t)(sizeof(f) / sizeof(f[0]));
	if (k > nmax) k = nmax;
	memcpy(out, f, k * sizeof(uint64_t));
	return (int)k;
}
