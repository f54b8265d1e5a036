/*
 * libecc_b200_dropin.h — the reference's OWN entry points for the hot path, on the reference's own structs,
 * implemented on the GPU engine (libecc_b200.h).  Built as libecc_b200/libecc_b200_dropin.so.
 *
 * What a libecc maintainer gets (see INTEGRATION.md for the exact link lines):
 *   - `prj_pt_mul` with the reference's exact signature (src/curves/prj_pt.h:61).  Linked (or LD_PRELOADed) ahead of
 *     a shared libsign/libec, it interposes the reference's definition, so every caller — ECDSA sign/verify
 *     (src/sig/ecdsa_common.c:479,788,793), ECC-CDH (src/ecdh/ecccdh.c:80,209), all other schemes of src/sig — runs
 *     its scalar multiplications on the B200 without being recompiled.
 *   - `ec_verify` with the reference's exact signature (src/sig/sig_algs.h:85-88): ECDSA / DECDSA / ECFSDSA / BIP0340
 *     verifications become ONE kernel launch each, the eight other short-Weierstrass schemes (ECKCDSA, ECSDSA, ECOSDSA,
 *     ECGDSA, ECRDSA, SM2, BIGN, DBIGN) one launch plus their host hashes; EdDSA is forwarded to the reference's own.
 *   - `prj_pt_mul_blind` (src/curves/prj_pt.h:62) is exported but NOT taken over by default: see its comment below.
 *   - `eccb200_dropin_prj_pt_mul_batch`: the same on arrays of reference structs (one launch for the batch).
 *   - `eccb200_dropin_ecdsa_verify_batch`: a function with the signature of the reference's per-scheme
 *     `verify_batch` slot (src/sig/sig_algs_internal.h:78-81), to put in ec_sig_maps[] where the reference has
 *     `unsupported_verify_batch` for ECDSA (:294).  Hashing stays in the reference's src/hash (resolved from the
 *     host application through get_hash_by_type, src/hash/hash_algs.h:550).
 *
 * The structs below MIRROR the reference's layout for its default configuration (WORDSIZE = 64,
 * NN_MAX_WORD_LEN = 27: src/nn/nn_config.h:154, complete formulas on).  They are declarations of an ABI, written for
 * this header; tests/test_dropin_layout.py checks every size and offset against the reference's own headers
 * (through oracle/ref_shim.c: ref_abi_facts) so that a configuration drift is caught.
 */
#ifndef LIBECC_B200_DROPIN_H
#define LIBECC_B200_DROPIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECCB200_NN_MAX_WORD_LEN 27

typedef uint64_t eccb200_word_t; /* word_t, src/words/words_64.h:25 */
typedef uint16_t eccb200_bitcnt_t; /* bitcnt_t, src/words/words.h:71 */

/* nn, src/nn/nn.h:67-71; magic = 0xb4cf5d56e2023316 ^ (NN_MAX_WORD_LEN + WORDSIZE)  (src/nn/nn.c:28) */
typedef struct {
	eccb200_word_t val[ECCB200_NN_MAX_WORD_LEN];
	eccb200_word_t magic;
	uint8_t wlen;
} eccb200_nn;

/* fp_ctx, src/fp/fp.h:31-57 */
typedef struct {
	eccb200_nn p;
	eccb200_bitcnt_t p_bitlen;
	eccb200_word_t mpinv;
	eccb200_nn r;
	eccb200_nn r_square;
	eccb200_bitcnt_t p_shift;
	eccb200_nn p_normalized;
	eccb200_word_t p_reciprocal;
	eccb200_word_t magic;
} eccb200_fp_ctx;

/* fp, src/fp/fp.h:73-77; magic 0x14e96c8ab28221ef (src/fp/fp.c:127) */
typedef struct {
	eccb200_nn fp_val;
	const eccb200_fp_ctx *ctx;
	eccb200_word_t magic;
} eccb200_fp;

/* ec_shortw_crv, src/curves/ec_shortw.h:25-36 (complete formulas enabled) */
typedef struct {
	eccb200_fp a;
	eccb200_fp b;
	eccb200_fp a_monty;
	eccb200_fp b3;
	eccb200_fp b_monty;
	eccb200_fp b3_monty;
	eccb200_nn order;
	eccb200_word_t magic;
} eccb200_ec_shortw_crv;

/* prj_pt, src/curves/prj_pt.h:26-32; magic 0xe1cd70babb1d5afe (src/curves/prj_pt.c:26) */
typedef struct {
	eccb200_fp X;
	eccb200_fp Y;
	eccb200_fp Z;
	const eccb200_ec_shortw_crv *crv;
	eccb200_word_t magic;
} eccb200_prj_pt;

/* head of ec_pub_key, src/sig/ec_key.h:120-131 (params is an opaque const ec_params *) */
typedef struct {
	int key_type; /* ec_alg_type */
	const void *params;
	eccb200_prj_pt y;
	eccb200_word_t magic;
} eccb200_ec_pub_key;

/*
 * Drop-in replacements with the reference's exact prototypes (src/curves/prj_pt.h:61-62):
 *     int prj_pt_mul(prj_pt_t out, nn_src_t m, prj_pt_src_t in);
 *     int prj_pt_mul_blind(prj_pt_t out, nn_src_t m, prj_pt_src_t in);
 * (prj_pt_mul_blind: see the security note below.)
 * Semantics kept (SURVEY.md §8a edge table): 0 / -1; `in` must be initialised and on its curve (else -1); any
 * scalar m (up to 27 words) gives (m mod order)*in; out may alias in; out is a valid initialised struct, canonical
 * (x, y, 1) for finite results and (0, 1, 0) for the point at infinity.  Supported curves: those of libecc_b200.h
 * (identified by p and the order); any other curve returns -1.
 */
int prj_pt_mul(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in);

/*
 * SECURITY NOTE — prj_pt_mul_blind.  The reference calls prj_pt_mul_blind where the scalar is SECRET (the ECDSA nonce
 * under USE_SIG_BLINDING, src/sig/ecdsa_common.c:476): it blinds the scalar and runs its constant-time ladder
 * (src/curves/prj_pt.c:1782-1822).  The GPU engine is a throughput path: table indices and branches depend on the
 * scalar, nothing is blinded.  Therefore this symbol does NOT silently route secret scalars to the GPU: by default it
 * forwards to the NEXT prj_pt_mul_blind in the process (the reference's own, when this library is preloaded or linked
 * ahead of a shared libec) and returns -1 if there is none.  Setting ECCB200_BLIND_ON_GPU=1 in the environment, or
 * calling eccb200_dropin_allow_nonct_blind(1), opts in to serving it on the GPU (same point, no side-channel
 * protection).  The same caveat holds for the explicit batch signing / ECDH entry points of libecc_b200.h, which a
 * caller only reaches by choosing them.
 */
int prj_pt_mul_blind(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in);
void eccb200_dropin_allow_nonct_blind(int on);

/*
 * ec_verify with the reference's prototype (src/sig/sig_algs.h:85-88; sig_type / hash_type are the reference's
 * ec_alg_type / hash_alg_type values):
 *     int ec_verify(const u8 *sig, u8 siglen, const ec_pub_key *pub_key, const u8 *m, u32 mlen,
 *                   ec_alg_type sig_type, hash_alg_type hash_type, const u8 *adata, u16 adata_len);
 * ECDSA (1), DECDSA (14), ECFSDSA (5) and BIP0340 (20) without ancillary data on a supported curve: the message is
 * hashed on the host with the reference's src/hash and the whole verification (steps 3-10 of __ecdsa_verify_finalize,
 * src/sig/ecdsa_common.c:760-810) is ONE launch of the verification kernel; 0 = valid, -1 = invalid.
 * ECKCDSA (2), ECSDSA (3), ECOSDSA (4), ECGDSA (6), ECRDSA (7), SM2 (8), BIGN (18) and DBIGN (19): the EC core
 * W' = a*G + b*Y is one launch of the double-scalar kernel, the scheme's mod-q scalar preparation, hashes and comparison
 * run on the host around it (SM2 and BIGN with their ancillary data: the signer's ID, the hash OID record).
 * Anything else - another scheme, an unknown curve, ancillary data on a scheme that takes none, SM2 / BIGN without
 * theirs - is forwarded unchanged to the next ec_verify in the process (the reference's own); -1 if there is none.
 */
int ec_verify(const uint8_t *sig, uint8_t siglen, const eccb200_ec_pub_key *pub_key, const uint8_t *m, uint32_t mlen,
	      int sig_type, int hash_type, const uint8_t *adata, uint16_t adata_len);
int eccb200_dropin_ec_verify(const uint8_t *sig, uint8_t siglen, const eccb200_ec_pub_key *pub_key, const uint8_t *m,
			     uint32_t mlen, int sig_type, int hash_type, const uint8_t *adata, uint16_t adata_len);

/*
 * ec_verify_batch and is_verify_batch_mode_supported with the reference's prototypes (src/sig/sig_algs.h:90-93,
 * src/sig/sig_algs.c:675-694, :937-958; scratch_pad_area is the reference's verify_batch_scratch_pad *, unused here):
 *     int ec_verify_batch(const u8 **s, const u8 *s_len, const ec_pub_key **pub_keys, const u8 **m, const u32 *m_len,
 *                         u32 num, ec_alg_type sig_type, hash_alg_type hash_type, const u8 **adata,
 *                         const u16 *adata_len, verify_batch_scratch_pad *scratch_pad_area, u32 *scratch_pad_area_len);
 *     int is_verify_batch_mode_supported(ec_alg_type sig_type, int *check);
 * The twelve schemes of the per-scheme adapters below (ECDSA, DECDSA, ECFSDSA, BIP0340, ECSDSA, ECOSDSA, ECKCDSA, ECGDSA,
 * ECRDSA, SM2, BIGN, DBIGN) are served by the device - ten of them sit at unsupported_verify_batch in the reference
 * (src/sig/sig_algs_internal.h:294); 0 iff ALL num signatures verify.  Other schemes (EdDSA), unknown curves and batches
 * with ancillary data on a scheme that takes none are forwarded to the next definition in the process (the reference's
 * own); -1 if there is none.
 */
int ec_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys, const uint8_t **m,
		    const uint32_t *m_len, uint32_t num, int sig_type, int hash_type, const uint8_t **adata,
		    const uint16_t *adata_len, void *scratch_pad_area, uint32_t *scratch_pad_area_len);
int eccb200_dropin_ec_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
				   const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
				   const uint8_t **adata, const uint16_t *adata_len, void *scratch_pad_area,
				   uint32_t *scratch_pad_area_len);
int is_verify_batch_mode_supported(int sig_type, int *check);

/* The same under non-clashing names (for callers that link the reference statically and choose per call). */
int eccb200_dropin_prj_pt_mul(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in);

/* Batched form on arrays of reference structs: out[i] = m[i] * in[i], all on the same curve; ret[i] (optional, may
 * be NULL) receives the per-item 0 / -1; returns 0 iff every item succeeded. */
int eccb200_dropin_prj_pt_mul_batch(eccb200_prj_pt *out, const eccb200_nn *m, const eccb200_prj_pt *in, uint32_t n,
				    int *ret);

/*
 * ECDSA back end for the reference's `verify_batch` slot (src/sig/sig_algs_internal.h:78-81; generic entry
 * ec_verify_batch src/sig/sig_algs.c:675).  Same contract as the other schemes' implementations
 * (e.g. src/sig/ecfsdsa.c:711-1074): returns 0 iff ALL num signatures verify, -1 otherwise; scratch_pad_area may be
 * NULL (it is not used); all keys must share one curve.  adata must be NULL (ECDSA takes none).
 */
int eccb200_dropin_ecdsa_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
				      const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type,
				      int hash_type, const uint8_t **adata, const uint16_t *adata_len,
				      void *scratch_pad_area, uint32_t *scratch_pad_area_len);

/* The same for ECFSDSA (sig_type = ECFSDSA = 5): a replacement for the reference's own ecfsdsa_verify_batch
 * (src/sig/ecfsdsa.c:1057) in the ECFSDSA entry of ec_sig_maps[]; signatures are W_x || W_y || s and the digest is
 * H(W_x || W_y || m), computed with the reference's src/hash. */
int eccb200_dropin_ecfsdsa_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
					const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type,
					int hash_type, const uint8_t **adata, const uint16_t *adata_len,
					void *scratch_pad_area, uint32_t *scratch_pad_area_len);

/* The same for BIP0340 (sig_type = BIP0340 = 20): a replacement for the reference's own bip0340_verify_batch
 * (src/sig/bip0340.c:1296) in the BIP0340 entry of ec_sig_maps[] (src/sig/sig_algs_internal.h:615); signatures are
 * r || s and the digest is the tagged challenge hash of r || x(Y) || m, computed with the reference's src/hash. */
int eccb200_dropin_bip0340_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
					const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type,
					int hash_type, const uint8_t **adata, const uint16_t *adata_len,
					void *scratch_pad_area, uint32_t *scratch_pad_area_len);

/* ECSDSA (sig_type 3) and ECOSDSA (4), whose entries the reference leaves at unsupported_verify_batch: checks and
 * e = -(r mod q) on the host, W' = sG + eY for the whole batch in one launch (eccb200_double_smul_batch), then
 * r' = H(W'x [|| W'y] || m) with the reference's src/hash (src/sig/ecsdsa_common.c:425-609). */
int eccb200_dropin_ecsdsa_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
				       const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type,
				       int hash_type, const uint8_t **adata, const uint16_t *adata_len,
				       void *scratch_pad_area, uint32_t *scratch_pad_area_len);

/* ECKCDSA (sig_type 2; src/sig/eckcdsa.c:543-832): h = H(z || m) and e = OS2I(r XOR h) mod q on the host, W' = sY + eG for
 * the whole batch in one launch, r' = H(W'_x) on the host. */
int eccb200_dropin_eckcdsa_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
					const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type,
					int hash_type, const uint8_t **adata, const uint16_t *adata_len,
					void *scratch_pad_area, uint32_t *scratch_pad_area_len);

/*
 * ECGDSA (6; src/sig/ecgdsa.c:413-600), ECRDSA (7; src/sig/ecrdsa.c:417-600), SM2 (8; src/sig/sm2.c:518-700) and
 * BIGN (18) / DBIGN (19; src/sig/bign_common.c:742-990), all left at unsupported_verify_batch by the reference.  Host:
 * range checks and the mod-q scalars - ECGDSA u = r^-1 e, v = r^-1 s; ECRDSA u = h^-1 s, v = -h^-1 r; SM2 (s, r + s);
 * BIGN (s1 + h, s0 + 2^(8l)) - with ONE inversion mod q per chunk of items (Montgomery's trick) where the reference does
 * one nn_modinv per signature; device: W' = u*G + v*Y for the whole batch in one launch; host: W'_x mod q == r
 * (ECGDSA, ECRDSA), (e + W'_x) mod q == r with e = H(Z || m) and Z built from adata[i] = the signer's ID (SM2),
 * BELT-HASH(OID || W' || H(m)) == s0 with the OID record in adata[i] (BIGN).  adata / adata_len are read per item for SM2
 * and BIGN (an item without its ancillary data is rejected, as the reference's ec_verify rejects it) and must be NULL
 * entries for ECGDSA / ECRDSA.  ECRDSA follows the reference's default build (digest byte-reversed, RFC flavour); set
 * ECCB200_ECRDSA_ISO14888_3=1 next to a libecc built with USE_ISO14888_3_ECRDSA (src/sig/ecrdsa.c:545-547).
 */
int eccb200_dropin_ecgdsa_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
				       const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type,
				       int hash_type, const uint8_t **adata, const uint16_t *adata_len,
				       void *scratch_pad_area, uint32_t *scratch_pad_area_len);
int eccb200_dropin_ecrdsa_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
				       const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type,
				       int hash_type, const uint8_t **adata, const uint16_t *adata_len,
				       void *scratch_pad_area, uint32_t *scratch_pad_area_len);
int eccb200_dropin_sm2_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
				    const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
				    const uint8_t **adata, const uint16_t *adata_len, void *scratch_pad_area,
				    uint32_t *scratch_pad_area_len);
int eccb200_dropin_bign_verify_batch(const uint8_t **s, const uint8_t *s_len, const eccb200_ec_pub_key **pub_keys,
				     const uint8_t **m, const uint32_t *m_len, uint32_t num, int sig_type, int hash_type,
				     const uint8_t **adata, const uint16_t *adata_len, void *scratch_pad_area,
				     uint32_t *scratch_pad_area_len);

/* Per-signature verdicts of the last eccb200_dropin_*_verify_batch call on this thread (0 / -1), for callers
 * that want to know WHICH signature failed; returns the number of verdicts copied. */
uint32_t eccb200_dropin_last_verdicts(int8_t *verdicts, uint32_t cap);

/* Number of scalar multiplications the drop-in layer has served in this process (lets a caller confirm that the
 * interposed prj_pt_mul, not the CPU one, was reached). */
unsigned long long eccb200_dropin_call_count(void);

/* Signatures verified on the GPU by this layer (ec_verify shim and verify_batch adapters). */
unsigned long long eccb200_dropin_verify_count(void);

/* ECFSDSA / BIP0340 batches of at least 16384 signatures (ECCB200_DROPIN_MSM_MIN) are first checked as ONE multi-scalar
 * multiplication (eccb200_*_verify_msm_batch, include/libecc_b200.h); a batch that passes is settled there, a batch that
 * fails goes through the per-item kernel to find the culprit.  This counts the batches settled by the fast path. */
unsigned long long eccb200_dropin_msm_batches(void);

/* Device the drop-in layer uses (default 0; also settable with the ECCB200_DEVICE environment variable). */
int eccb200_dropin_set_device(int device);

/*
 * Memory policy.  Engine contexts are created lazily, per curve: up to four SMALL ones (16-bit fixed-base comb,
 * 64 MiB for a 256-bit curve; concurrent callers do not block each other) for single calls and small batches, and
 * one BIG one (the engine's default 22-bit comb, 3.2 GiB for a 256-bit curve; ECCB200_COMB_WINDOW overrides) only
 * when a batch of at least 2^15 items arrives.  eccb200_dropin_release() destroys every context and frees the
 * page-locked staging buffers; the next call re-creates what it needs.
 */
void eccb200_dropin_release(void);

#ifdef __cplusplus
}
#endif
#endif
