/*
 * oracle/ecc_oracle.h — CPU restatement ("port") of the reference's prj_pt_mul / ECDSA-verify path.
 *
 * TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * are the only permitted callers.  The product (libecc_b200/) never links, loads or calls this.
 *
 * Parity status: PINNED — tests/test_oracle.py checks this port against (i) the golden vectors extracted from
 * the reference's own test headers (the tests/golden JSON files: NIST ECC-CDH, RFC 4754/6979 ECDSA, Wycheproof) and
 * (ii) the unmodified reference compiled here (oracle/_ref/libecc_ref.so) on seeded random inputs.
 */
#ifndef ECC_ORACLE_H
#define ECC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* curve: "SECP256R1" | "FRP256V1" | "SECP384R1".  All byte strings big-endian (libecc wire format). */

/* out[i] = affine (x||y) of scalars[i] * (points ? points[i] : G); status 0 finite, 1 infinity, -1 error. */
int ora_prj_pt_mul_batch(const char *curve, uint32_t n, const uint8_t *scalars, uint32_t slen,
			 const uint8_t *points, uint8_t *out, int8_t *status, int nthreads);

/* ECDSA verification on pre-hashed messages; digests[i] has hlen bytes.  verdict 0 valid, -1 invalid/error. */
int ora_ecdsa_verify_digest_batch(const char *curve, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				  const uint8_t *digests, uint32_t hlen, int8_t *verdict, int nthreads);

/* ECFSDSA verification (sig/ecfsdsa.c) on digests[i] = H(r_i || m_i); sigs are [n][2*plen + qlen] (r = W_x || W_y, s). */
int ora_ecfsdsa_verify_digest_batch(const char *curve, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				    const uint8_t *digests, uint32_t hlen, int8_t *verdict, int nthreads);

/* W_i = a_i*G + b_i*Y_i with affine results — the EC core of the reference's Schnorr-type verifications
 * (sig/ecsdsa_common.c:493-497 ...); ab [n][2*qlen]; status 0 finite / 1 infinity / -1 key rejected. */
int ora_double_smul_batch(const char *curve, uint32_t n, const uint8_t *ab, const uint8_t *pubkeys, uint8_t *out,
			  int8_t *status, int nthreads);

/* BIP0340 verification (sig/bip0340.c:383-577) on digests[i] = H(H(tag) || H(tag) || r_i || x(Y_i) || m_i) with
 * tag = "BIP0340/challenge"; sigs are [n][plen + qlen] (r = x coordinate, s). */
int ora_bip0340_verify_digest_batch(const char *curve, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				    const uint8_t *digests, uint32_t hlen, int8_t *verdict, int nthreads);

/* ECDSA signing on pre-hashed messages with caller-supplied nonces (deterministic; RFC-vector friendly).
 * status 0 ok, -1 error (k, d out of range, r == 0 or s == 0). */
int ora_ecdsa_sign_digest_batch(const char *curve, uint32_t n, const uint8_t *privkeys, const uint8_t *nonces,
				const uint8_t *digests, uint32_t hlen, uint8_t *sigs, int8_t *status, int nthreads);

/* out = a*b*R^-1 mod p with R = 2^(64*nlimbs) (Montgomery product on raw values < p), plen bytes each. */
int ora_fp_mul_monty(const char *curve, const uint8_t *a, const uint8_t *b, uint8_t *out);

/* Number of Montgomery multiplications executed by the last ora_prj_pt_mul_batch call on this thread with n == 1
 * (the reference algorithm's M_ref). */
uint64_t ora_last_mul_count(void);

int ora_curve_sizes(const char *curve, uint32_t *plen, uint32_t *qlen);

#ifdef __cplusplus
}
#endif
#endif
