/*
 * tests/golden/dump_golden.c — extracts the reference's own known-answer vectors for the prj_pt_mul / ECDSA
 * path into neutral JSON fixtures (tests/golden/*.json).  Run by tests/golden/make_golden.sh in the build
 * container, where /root/reference exists; the fixtures (not this program's inputs) travel to the GPU box.
 *
 * It #includes the reference's test-vector headers in place (nothing is copied into the repo) and links
 * oracle/_ref/libecc_ref.so, so every "ref_*" field below is an output of the unmodified reference run here.
 *
 * Sources (relative to /root/reference/src):
 *   tests/ec_self_tests_core.h  ec_fixed_vector_tests[] (:4915)  — ECDSA / DECDSA KATs (RFC 4754, RFC 6979, ...)
 *                               ecdh_fixed_vector_tests[] (:5294) — NIST ECC-CDH KATs
 *   wycheproof_tests/libecc_wycheproof_tests.h  wycheproof_ecdsa_all_tests[] (:329861),
 *                               wycheproof_ecdh_all_tests[] (:539273)
 */
#include "libsig.h"
#include "tests/ec_self_tests_core.h"
#include "wycheproof_tests/libecc_wycheproof.h"
#include "wycheproof_tests/libecc_wycheproof_tests.h"
#include <stdio.h>
#include <string.h>

static void hex(FILE *f, const char *key, const u8 *b, unsigned int len, int last)
{
	fprintf(f, "\"%s\": \"", key);
	for (unsigned int i = 0; i < len; i++) fprintf(f, "%02x", b[i]);
	fprintf(f, "\"%s", last ? "" : ", ");
}

static const char *curve_name(const ec_str_params *sp)
{
	return (const char *)sp->name->buf;
}

static int wanted_curve(const ec_str_params *sp)
{
	const char *n = curve_name(sp);
	return !strcmp(n, "SECP256R1") || !strcmp(n, "SECP384R1") || !strcmp(n, "FRP256V1") ||
	       !strcmp(n, "BRAINPOOLP256R1") || !strcmp(n, "BRAINPOOLP384R1") || !strcmp(n, "SECP256K1") ||
	       !strcmp(n, "SECP521R1") || !strcmp(n, "SM2P256V1") || !strcmp(n, "BRAINPOOLP512R1") ||
	       !strcmp(n, "SECP224R1") || !strcmp(n, "SECP192R1");
}

static const char *hash_name(hash_alg_type t)
{
	const hash_mapping *hm = NULL;
	if (get_hash_by_type(t, &hm) || !hm) return "?";
	return hm->name;
}

static void jstr(FILE *f, const char *key, const char *s, int last)
{
	fprintf(f, "\"%s\": \"", key);
	for (; s && *s; s++) {
		if (*s == '"' || *s == '\\') fputc('\\', f);
		if ((unsigned char)*s >= 0x20) fputc(*s, f);
	}
	fprintf(f, "\"%s", last ? "" : ", ");
}

static int digest_of(hash_alg_type t, const u8 *m, u32 mlen, u8 *out, u8 *dlen)
{
	const hash_mapping *hm = NULL;
	const u8 *in[2] = { m, NULL };
	u32 il[1] = { mlen };
	if (get_hash_by_type(t, &hm) || !hm) return -1;
	*dlen = hm->digest_size;
	return hm->hfunc_scattered(in, il, out);
}

int main(int argc, char **argv)
{
	const char *dir = (argc > 1) ? argv[1] : ".";
	char path[512];
	FILE *f;
	int first;

	/* ---------------------------------------------------------------- ECC-CDH */
	snprintf(path, sizeof(path), "%s/ecccdh_kat.json", dir);
	f = fopen(path, "w");
	fprintf(f, "[\n");
	first = 1;
	for (unsigned int i = 0; i < sizeof(ecdh_fixed_vector_tests) / sizeof(ecdh_fixed_vector_tests[0]); i++) {
		const ecdh_test_case *t = ecdh_fixed_vector_tests[i];
		if (!t || t->ecdh_type != ECCCDH || !wanted_curve(t->ec_str_p)) continue;
		fprintf(f, "%s {", first ? "" : ",\n");
		first = 0;
		jstr(f, "name", t->name, 0);
		jstr(f, "curve", curve_name(t->ec_str_p), 0);
		hex(f, "priv", t->our_priv_key, t->our_priv_key_len, 0);
		hex(f, "peer_pub", t->peer_pub_key, t->peer_pub_key_len, 0);
		hex(f, "our_pub", t->exp_our_pub_key, t->exp_our_pub_key_len, 0);
		hex(f, "shared", t->exp_shared_secret, t->exp_shared_secret_len, 1);
		fprintf(f, "}");
	}
	fprintf(f, "\n]\n");
	fclose(f);

	/* ---------------------------------------------------------------- ECDSA / DECDSA KATs */
	snprintf(path, sizeof(path), "%s/ecdsa_kat.json", dir);
	f = fopen(path, "w");
	fprintf(f, "[\n");
	first = 1;
	for (unsigned int i = 0; i < sizeof(ec_fixed_vector_tests) / sizeof(ec_fixed_vector_tests[0]); i++) {
		const ec_test_case *t = ec_fixed_vector_tests[i];
		ec_params params;
		ec_key_pair kp;
		u8 pub[2 * 66], dg[MAX_DIGEST_SIZE], dlen = 0, kbuf[66];
		u8 plen, qlen;
		int ref_verdict, have_k = 0;
		if (!t || (t->sig_type != ECDSA && t->sig_type != DECDSA) || !wanted_curve(t->ec_str_p)) continue;
		if (import_params(&params, t->ec_str_p)) return 1;
		plen = (u8)BYTECEIL(params.ec_fp.p_bitlen);
		qlen = (u8)BYTECEIL(params.ec_gen_order_bitlen);
		if (ec_key_pair_import_from_priv_key_buf(&kp, &params, t->priv_key, t->priv_key_len, t->sig_type)) return 1;
		if (ec_pub_key_export_to_aff_buf(&kp.pub_key, pub, (u8)(2 * plen))) return 1;
		if (digest_of(t->hash_type, (const u8 *)t->msg, t->msglen, dg, &dlen)) return 1;
		ref_verdict = ec_verify(t->exp_sig, t->exp_siglen, &kp.pub_key, (const u8 *)t->msg, t->msglen,
					t->sig_type, t->hash_type, t->adata, t->adata_len);
		if (t->nn_random) { /* the nonce the reference's harness injects (ec_self_tests_core.h:34) */
			nn k;
			if (!t->nn_random(&k, &params.ec_gen_order) && !nn_export_to_buf(kbuf, qlen, &k)) have_k = 1;
		}
		fprintf(f, "%s {", first ? "" : ",\n");
		first = 0;
		jstr(f, "name", t->name, 0);
		jstr(f, "curve", curve_name(t->ec_str_p), 0);
		jstr(f, "alg", t->sig_type == ECDSA ? "ECDSA" : "DECDSA", 0);
		jstr(f, "hash", hash_name(t->hash_type), 0);
		hex(f, "priv", t->priv_key, t->priv_key_len, 0);
		hex(f, "pub", pub, (unsigned int)(2 * plen), 0);
		hex(f, "msg", (const u8 *)t->msg, t->msglen, 0);
		hex(f, "digest", dg, dlen, 0);
		if (have_k) hex(f, "nonce", kbuf, qlen, 0);
		hex(f, "sig", t->exp_sig, t->exp_siglen, 0);
		fprintf(f, "\"ref_verdict\": %d}", ref_verdict);
	}
	fprintf(f, "\n]\n");
	fclose(f);

	/* ---------------------------------------------------------------- ECFSDSA KATs (sig/ecfsdsa.c) */
	snprintf(path, sizeof(path), "%s/ecfsdsa_kat.json", dir);
	f = fopen(path, "w");
	fprintf(f, "[\n");
	first = 1;
	for (unsigned int i = 0; i < sizeof(ec_fixed_vector_tests) / sizeof(ec_fixed_vector_tests[0]); i++) {
		const ec_test_case *t = ec_fixed_vector_tests[i];
		ec_params params;
		ec_key_pair kp;
		u8 pub[2 * 66], dg[MAX_DIGEST_SIZE], plen;
		const hash_mapping *hm = NULL;
		const u8 *in[3];
		u32 il[2];
		int ref_verdict;
		if (!t || t->sig_type != ECFSDSA || !wanted_curve(t->ec_str_p)) continue;
		if (import_params(&params, t->ec_str_p)) return 1;
		plen = (u8)BYTECEIL(params.ec_fp.p_bitlen);
		if (ec_key_pair_import_from_priv_key_buf(&kp, &params, t->priv_key, t->priv_key_len, t->sig_type)) return 1;
		if (ec_pub_key_export_to_aff_buf(&kp.pub_key, pub, (u8)(2 * plen))) return 1;
		if (get_hash_by_type(t->hash_type, &hm) || !hm) return 1;
		/* h = H(r || m): the first 2*plen bytes of the signature, then the message (sig/ecfsdsa.c:482,529) */
		in[0] = t->exp_sig; il[0] = (u32)(2 * plen);
		in[1] = (const u8 *)t->msg; il[1] = t->msglen;
		in[2] = NULL;
		if (hm->hfunc_scattered(in, il, dg)) return 1;
		ref_verdict = ec_verify(t->exp_sig, t->exp_siglen, &kp.pub_key, (const u8 *)t->msg, t->msglen,
					t->sig_type, t->hash_type, t->adata, t->adata_len);
		fprintf(f, "%s {", first ? "" : ",\n");
		first = 0;
		jstr(f, "name", t->name, 0);
		jstr(f, "curve", curve_name(t->ec_str_p), 0);
		jstr(f, "hash", hash_name(t->hash_type), 0);
		hex(f, "priv", t->priv_key, t->priv_key_len, 0);
		hex(f, "pub", pub, (unsigned int)(2 * plen), 0);
		hex(f, "msg", (const u8 *)t->msg, t->msglen, 0);
		hex(f, "digest_rm", dg, hm->digest_size, 0);
		hex(f, "sig", t->exp_sig, t->exp_siglen, 0);
		fprintf(f, "\"ref_verdict\": %d}", ref_verdict);
	}
	fprintf(f, "\n]\n");
	fclose(f);

	/* ---------------------------------------------------------------- BIP0340 KATs (sig/bip0340.c) */
	snprintf(path, sizeof(path), "%s/bip0340_kat.json", dir);
	f = fopen(path, "w");
	fprintf(f, "[\n");
	first = 1;
	for (unsigned int i = 0; i < sizeof(ec_fixed_vector_tests) / sizeof(ec_fixed_vector_tests[0]); i++) {
		const ec_test_case *t = ec_fixed_vector_tests[i];
		ec_params params;
		ec_key_pair kp;
		u8 pub[2 * 66], dg[MAX_DIGEST_SIZE], htag[MAX_DIGEST_SIZE], plen;
		const hash_mapping *hm = NULL;
		const u8 *in[6];
		u32 il[5];
		int ref_verdict;
		static const char tag[] = "BIP0340/challenge";
		if (!t || t->sig_type != BIP0340 || !wanted_curve(t->ec_str_p)) continue;
		if (import_params(&params, t->ec_str_p)) return 1;
		plen = (u8)BYTECEIL(params.ec_fp.p_bitlen);
		if (ec_key_pair_import_from_priv_key_buf(&kp, &params, t->priv_key, t->priv_key_len, t->sig_type)) return 1;
		if (ec_pub_key_export_to_aff_buf(&kp.pub_key, pub, (u8)(2 * plen))) return 1;
		if (get_hash_by_type(t->hash_type, &hm) || !hm) return 1;
		/* the challenge hash: H(H(tag) || H(tag) || r || x(Y) || m)  (sig/bip0340.c:45-69, :438-443) */
		in[0] = (const u8 *)tag; il[0] = (u32)(sizeof(tag) - 1);
		in[1] = NULL;
		if (hm->hfunc_scattered(in, il, htag)) return 1;
		in[0] = htag; il[0] = hm->digest_size;
		in[1] = htag; il[1] = hm->digest_size;
		in[2] = t->exp_sig; il[2] = plen;
		in[3] = pub; il[3] = plen;
		in[4] = (const u8 *)t->msg; il[4] = t->msglen;
		in[5] = NULL;
		if (hm->hfunc_scattered(in, il, dg)) return 1;
		ref_verdict = ec_verify(t->exp_sig, t->exp_siglen, &kp.pub_key, (const u8 *)t->msg, t->msglen,
					t->sig_type, t->hash_type, t->adata, t->adata_len);
		fprintf(f, "%s {", first ? "" : ",\n");
		first = 0;
		jstr(f, "name", t->name, 0);
		jstr(f, "curve", curve_name(t->ec_str_p), 0);
		jstr(f, "hash", hash_name(t->hash_type), 0);
		hex(f, "priv", t->priv_key, t->priv_key_len, 0);
		hex(f, "pub", pub, (unsigned int)(2 * plen), 0);
		hex(f, "msg", (const u8 *)t->msg, t->msglen, 0);
		hex(f, "digest_challenge", dg, hm->digest_size, 0);
		hex(f, "sig", t->exp_sig, t->exp_siglen, 0);
		fprintf(f, "\"ref_verdict\": %d}", ref_verdict);
	}
	fprintf(f, "\n]\n");
	fclose(f);

	/* ---------------------------------------------------------------- Wycheproof ECDSA */
	snprintf(path, sizeof(path), "%s/wycheproof_ecdsa.json", dir);
	f = fopen(path, "w");
	fprintf(f, "[\n");
	first = 1;
	for (unsigned int i = 0; i < NUM_WYCHEPROOF_ECDSA_TESTS; i++) {
		const wycheproof_ecdsa_test *t = wycheproof_ecdsa_all_tests[i];
		ec_params params;
		ec_pub_key pk;
		u8 dg[MAX_DIGEST_SIZE], dlen = 0;
		int import_ret, ref_verdict;
		if (!t || t->sig_alg != ECDSA || !wanted_curve(t->curve)) continue;
		if (import_params(&params, t->curve)) return 1;
		if (digest_of(t->hash, t->msg, t->msglen, dg, &dlen)) return 1;
		import_ret = ec_pub_key_import_from_aff_buf(&pk, &params, t->pubkey, (u8)t->pubkeylen, t->sig_alg);
		ref_verdict = -1;
		if (!import_ret)
			ref_verdict = ec_verify(t->sig, (u8)t->siglen, &pk, t->msg, t->msglen, t->sig_alg, t->hash, NULL, 0);
		fprintf(f, "%s {", first ? "" : ",\n");
		first = 0;
		jstr(f, "name", t->name, 0);
		jstr(f, "curve", curve_name(t->curve), 0);
		jstr(f, "hash", hash_name(t->hash), 0);
		hex(f, "pub", t->pubkey, t->pubkeylen, 0);
		hex(f, "digest", dg, dlen, 0);
		hex(f, "sig", t->sig, t->siglen, 0);
		jstr(f, "comment", t->comment, 0);
		fprintf(f, "\"expected\": %d, \"ref_import\": %d, \"ref_verdict\": %d}", t->result, import_ret, ref_verdict);
	}
	fprintf(f, "\n]\n");
	fclose(f);

	/* ---------------------------------------------------------------- Wycheproof ECDH (uncompressed peers) */
	snprintf(path, sizeof(path), "%s/wycheproof_ecdh.json", dir);
	f = fopen(path, "w");
	fprintf(f, "[\n");
	first = 1;
	for (unsigned int i = 0; i < NUM_WYCHEPROOF_ECDH_TESTS; i++) {
		const wycheproof_ecdh_test *t = wycheproof_ecdh_all_tests[i];
		ec_params params;
		prj_pt Q, S;
		nn d;
		u8 plen, out[2 * 66];
		int ref_status = -1, iszero = 0;
		if (!t || t->ecdh_alg != ECCCDH || !wanted_curve(t->curve) || t->compressed) continue;
		if (import_params(&params, t->curve)) return 1;
		plen = (u8)BYTECEIL(params.ec_fp.p_bitlen);
		if (t->peerpubkeylen != (unsigned int)(2 * plen)) continue;
		memset(out, 0, sizeof(out));
		if (!nn_init_from_buf(&d, t->privkey, (u16)t->privkeylen) &&
		    !prj_pt_import_from_aff_buf(&Q, t->peerpubkey, (u16)t->peerpubkeylen, &params.ec_curve) &&
		    !prj_pt_mul(&S, &d, &Q) && !prj_pt_iszero(&S, &iszero)) {
			if (iszero) ref_status = 1;
			else if (!prj_pt_export_to_aff_buf(&S, out, (u32)(2 * plen))) ref_status = 0;
		}
		fprintf(f, "%s {", first ? "" : ",\n");
		first = 0;
		jstr(f, "name", t->name, 0);
		jstr(f, "curve", curve_name(t->curve), 0);
		hex(f, "priv", t->privkey, t->privkeylen, 0);
		hex(f, "peer_pub", t->peerpubkey, t->peerpubkeylen, 0);
		hex(f, "shared", t->sharedsecret, t->sharedsecretlen, 0);
		hex(f, "ref_point", out, (unsigned int)(2 * plen), 0);
		jstr(f, "comment", t->comment, 0);
		fprintf(f, "\"expected\": %d, \"ref_status\": %d}", t->result, ref_status);
	}
	fprintf(f, "\n]\n");
	fclose(f);
	return 0;
}
