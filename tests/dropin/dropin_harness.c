/*
 * tests/dropin/dropin_harness.c — exercises include/libecc_b200_dropin.h against the UNMODIFIED reference.
 *
 * Compiled in the build container against the reference's own headers (-I/root/reference/src) and linked with
 * oracle/_ref/libecc_ref.so, so every struct passed to the drop-in is a REAL reference struct (import_params,
 * prj_pt_import_from_aff_buf, nn_init_from_buf, ec_key_pair_gen ...) and every result is judged by the reference's own
 * predicates (prj_pt_check_initialized, prj_pt_is_on_curve, prj_pt_cmp, prj_pt_iszero, ec_verify).
 * The binary (oracle/_ref/dropin_harness) travels to the GPU box; tests/test_gpu_dropin.py runs it there.
 *
 * Modes:
 *   direct   the drop-in is dlopen'ed privately; its prj_pt_mul / batch / verify_batch entry points are compared with
 *            the reference's own functions on identical inputs.
 *   kats     every fixed-vector case of the reference's own self tests (src/tests/ec_self_tests_core.h) through the
 *            drop-in's ec_verify and a one-item ec_verify_batch, judged by the reference's ec_verify.
 *   fuzzmul  prj_pt_mul on random scalars of every nn width and on projective / infinite / off-curve / aliased points.
 *   fuzz     mutated signatures / keys / ancillary data of every served scheme: drop-in verdict == reference verdict.
 *   preload  run with LD_PRELOAD=libecc_b200_dropin.so: the reference's ec_sign / ec_verify / ECC-CDH code then
 *            calls the interposed prj_pt_mul, i.e. the GPU, without being recompiled; results must still satisfy
 *            the reference's known-answer expectations.
 */
#define _GNU_SOURCE
#include "libsig.h"
#include "tests/ec_self_tests_core.h" /* the reference's own known-answer vectors (src/tests): ec_fixed_vector_tests[] */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int (*mul_fn)(prj_pt_t, nn_src_t, prj_pt_src_t);
typedef int (*mul_batch_fn)(prj_pt *, const nn *, const prj_pt *, u32, int *);
typedef int (*vbatch_fn)(const u8 **, const u8 *, const ec_pub_key **, const u8 **, const u32 *, u32, ec_alg_type,
			 hash_alg_type, const u8 **, const u16 *, verify_batch_scratch_pad *, u32 *);
typedef u32 (*verdicts_fn)(signed char *, u32);
typedef unsigned long long (*count_fn)(void);
typedef void (*allow_fn)(int);
typedef int (*everify_fn)(const u8 *, u8, const ec_pub_key *, const u8 *, u32, ec_alg_type, hash_alg_type, const u8 *, u16);
#include <pthread.h>
#include <time.h>

static int failures = 0;
#define CHECK(cond, ...)                                      \
	do {                                                  \
		if (!(cond)) {                                \
			failures++;                           \
			printf("FAIL %s:%d: ", __FILE__, __LINE__); \
			printf(__VA_ARGS__);                  \
			printf("\n");                         \
		}                                             \
	} while (0)

static unsigned long long rng_state = 0x6c69626563632d31ULL;
static u8 rnd8(void)
{
	rng_state += 0x9e3779b97f4a7c15ULL;
	unsigned long long z = rng_state;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return (u8)((z ^ (z >> 31)) >> 24);
}

static int load_params(ec_params *params, const char *name)
{
	const ec_str_params *sp = NULL;
	if (ec_get_curve_params_by_name((const u8 *)name, (u8)(strlen(name) + 1), &sp) || !sp) return -1;
	return import_params(params, sp);
}

static void compare(const char *what, const char *curve, int r_ref, prj_pt_src_t o_ref, int r_gpu, prj_pt_src_t o_gpu)
{
	int z1 = 0, z2 = 0, cmp = 1, on = 0;
	CHECK(r_ref == r_gpu, "%s %s: return %d vs reference %d", curve, what, r_gpu, r_ref);
	if (r_ref || r_gpu) return;
	CHECK(!prj_pt_check_initialized(o_gpu), "%s %s: output not a valid initialised prj_pt", curve, what);
	CHECK(!prj_pt_is_on_curve(o_gpu, &on) && on, "%s %s: output fails the reference's on-curve check", curve, what);
	CHECK(!prj_pt_iszero(o_ref, &z1) && !prj_pt_iszero(o_gpu, &z2) && z1 == z2, "%s %s: infinity flag", curve, what);
	if (!z1 && !z2) CHECK(!prj_pt_cmp(o_ref, o_gpu, &cmp) && cmp == 0, "%s %s: prj_pt_cmp != 0", curve, what);
}

static int run_direct(const char *dropin_path)
{
	void *h = dlopen(dropin_path, RTLD_NOW | RTLD_LOCAL);
	if (!h) {
		printf("FAIL dlopen %s: %s\n", dropin_path, dlerror());
		return 1;
	}
	mul_fn gpu_mul = (mul_fn)dlsym(h, "prj_pt_mul");
	mul_fn gpu_mul_blind = (mul_fn)dlsym(h, "prj_pt_mul_blind");
	mul_batch_fn gpu_batch = (mul_batch_fn)dlsym(h, "eccb200_dropin_prj_pt_mul_batch");
	vbatch_fn gpu_vbatch = (vbatch_fn)dlsym(h, "eccb200_dropin_ecdsa_verify_batch");
	vbatch_fn gpu_fsbatch = (vbatch_fn)dlsym(h, "eccb200_dropin_ecfsdsa_verify_batch");
	verdicts_fn gpu_verdicts = (verdicts_fn)dlsym(h, "eccb200_dropin_last_verdicts");
	count_fn gpu_calls = (count_fn)dlsym(h, "eccb200_dropin_call_count");
	count_fn gpu_vcount = (count_fn)dlsym(h, "eccb200_dropin_verify_count");
	allow_fn gpu_allow_blind = (allow_fn)dlsym(h, "eccb200_dropin_allow_nonct_blind");
	everify_fn gpu_everify = (everify_fn)dlsym(h, "eccb200_dropin_ec_verify");
	if (!gpu_mul || !gpu_mul_blind || !gpu_batch || !gpu_vbatch || !gpu_fsbatch || !gpu_verdicts || !gpu_calls ||
	    !gpu_vcount || !gpu_allow_blind || !gpu_everify) {
		printf("FAIL missing drop-in symbols\n");
		return 1;
	}
	const char *names[8] = { "SECP256R1", "FRP256V1", "SECP384R1", "BRAINPOOLP256R1", "SECP256K1", "SECP521R1",
				 "SECP224R1", "BRAINPOOLP512R1" };
	/* a scheme on a curve the layer does not know (EdDSA on WEI25519) is forwarded to the reference's own ec_verify */
	{
		ec_params wp;
		ec_key_pair kq;
		u8 sg[2 * 66], sgl = 0;
		const u8 msg[5] = { 'h', 'e', 'l', 'l', 'o' };
		if (!load_params(&wp, "WEI25519") && !ec_key_pair_gen(&kq, &wp, EDDSA25519) &&
		    !ec_get_sig_len(&wp, EDDSA25519, SHA512, &sgl) && !ec_sign(sg, sgl, &kq, msg, 5, EDDSA25519, SHA512, NULL, 0)) {
			unsigned long long v0 = gpu_vcount();
			CHECK(gpu_everify(sg, sgl, &kq.pub_key, msg, 5, EDDSA25519, SHA512, NULL, 0) == 0, "forwarded EdDSA verify failed");
			sg[2] ^= 1;
			CHECK(gpu_everify(sg, sgl, &kq.pub_key, msg, 5, EDDSA25519, SHA512, NULL, 0) == -1, "forwarded EdDSA forgery accepted");
			CHECK(gpu_vcount() == v0, "EdDSA must be forwarded, not counted as a GPU verification");
		} else {
			printf("note: the reference does not sign EDDSA25519 here; forwarding check skipped\n");
		}
	}
	const char *only = getenv("HARNESS_CURVES"); /* optional comma-separated subset (the CPU run of the tests) */
	/* HARNESS_SECTIONS: optional subset of the sections below, letters m (multiplications), e (ECDSA), f (ECFSDSA),
	 * s (ECSDSA / ECOSDSA / ECKCDSA), d (ECGDSA / ECRDSA / SM2 / BIGN / DBIGN), b (BIP0340); default: all */
	const char *sect = getenv("HARNESS_SECTIONS");
#define SECTION(ch) (!sect || strchr(sect, (ch)))
	for (int c = 0; c < 8; c++) {
		ec_params params;
		if (only && !strstr(only, names[c])) continue;
		CHECK(!load_params(&params, names[c]), "import_params %s", names[c]);
		u8 qlen = (u8)BYTECEIL(params.ec_gen_order_bitlen);
		/* ---- scalars of several widths on G, on a reference-made projective point (Z != 1), aliasing */
		prj_pt base2, tmp;
		CHECK(!prj_pt_dbl(&tmp, &params.ec_gen), "dbl");
		CHECK(!prj_pt_add(&base2, &tmp, &params.ec_gen), "add"); /* 3G with Z != 1 */
		u16 lens[5] = { qlen, 1, (u16)(qlen + 8), 72, (u16)(2 * qlen) };
		for (int t = 0; t < 10 && SECTION('m'); t++) {
			u8 kb[160];
			u16 kl = lens[t % 5];
			for (u16 i = 0; i < kl; i++) kb[i] = rnd8();
			nn k;
			CHECK(!nn_init_from_buf(&k, kb, kl), "nn_init_from_buf");
			prj_pt o_ref, o_gpu;
			prj_pt_src_t b = (t & 1) ? &base2 : &params.ec_gen;
			int r1 = prj_pt_mul(&o_ref, &k, b);
			int r2 = gpu_mul(&o_gpu, &k, b);
			compare("prj_pt_mul", names[c], r1, &o_ref, r2, &o_gpu);
			/* prj_pt_mul_blind: by default NOT served by the GPU (secret scalars) but forwarded to the reference's
			 * own; after the explicit opt-in it runs on the GPU.  Same point either way. */
			unsigned long long before = gpu_calls();
			r2 = gpu_mul_blind(&o_gpu, &k, b);
			compare("prj_pt_mul_blind (forwarded)", names[c], r1, &o_ref, r2, &o_gpu);
			CHECK(gpu_calls() == before, "%s: prj_pt_mul_blind reached the GPU without the opt-in", names[c]);
			gpu_allow_blind(1);
			r2 = gpu_mul_blind(&o_gpu, &k, b);
			compare("prj_pt_mul_blind (opt-in, GPU)", names[c], r1, &o_ref, r2, &o_gpu);
			CHECK(gpu_calls() == before + 1, "%s: opted-in prj_pt_mul_blind did not reach the GPU", names[c]);
			gpu_allow_blind(0);
			/* out == in */
			prj_pt alias;
			CHECK(!prj_pt_copy(&alias, b), "copy");
			r2 = gpu_mul(&alias, &k, &alias);
			compare("prj_pt_mul(out==in)", names[c], r1, &o_ref, r2, &alias);
		}
		/* ---- edge scalars: 0, 1, q-1, q, q+1 */
		if (SECTION('m')) {
			nn q, one, k;
			prj_pt o_ref, o_gpu;
			CHECK(!nn_copy(&q, &params.ec_gen_order) && !nn_init(&one, 0) && !nn_one(&one), "nn setup");
			CHECK(!nn_init(&k, 0) && !nn_zero(&k), "zero");
			compare("k=0", names[c], prj_pt_mul(&o_ref, &k, &params.ec_gen), &o_ref, gpu_mul(&o_gpu, &k, &params.ec_gen), &o_gpu);
			compare("k=1", names[c], prj_pt_mul(&o_ref, &one, &params.ec_gen), &o_ref, gpu_mul(&o_gpu, &one, &params.ec_gen), &o_gpu);
			compare("k=q", names[c], prj_pt_mul(&o_ref, &q, &params.ec_gen), &o_ref, gpu_mul(&o_gpu, &q, &params.ec_gen), &o_gpu);
			CHECK(!nn_sub(&k, &q, &one), "q-1");
			compare("k=q-1", names[c], prj_pt_mul(&o_ref, &k, &base2), &o_ref, gpu_mul(&o_gpu, &k, &base2), &o_gpu);
			CHECK(!nn_add(&k, &q, &one), "q+1");
			compare("k=q+1", names[c], prj_pt_mul(&o_ref, &k, &base2), &o_ref, gpu_mul(&o_gpu, &k, &base2), &o_gpu);
			/* in = infinity */
			prj_pt inf;
			CHECK(!prj_pt_init(&inf, &params.ec_curve) && !prj_pt_zero(&inf), "inf");
			compare("in=inf", names[c], prj_pt_mul(&o_ref, &one, &inf), &o_ref, gpu_mul(&o_gpu, &one, &inf), &o_gpu);
			/* point not on the curve: both must fail */
			prj_pt bad;
			CHECK(!prj_pt_copy(&bad, &params.ec_gen), "copy");
			bad.Y.fp_val.val[0] ^= 1;
			compare("off-curve", names[c], prj_pt_mul(&o_ref, &one, &bad), &o_ref, gpu_mul(&o_gpu, &one, &bad), &o_gpu);
			/* uninitialised input */
			prj_pt junk;
			memset(&junk, 0, sizeof(junk));
			CHECK(gpu_mul(&o_gpu, &one, &junk) == -1, "%s: uninitialised input accepted", names[c]);
		}
		/* ---- batch on arrays of structs */
		if (SECTION('m')) {
			enum { NB = 64 };
			static prj_pt in[NB], out[NB], ref_out[NB];
			static nn ks[NB];
			int rets[NB];
			for (int i = 0; i < NB; i++) {
				u8 kb[96];
				for (int j = 0; j < qlen; j++) kb[j] = rnd8();
				CHECK(!nn_init_from_buf(&ks[i], kb, qlen), "nn");
				CHECK(!prj_pt_copy(&in[i], (i % 3) ? &base2 : &params.ec_gen), "copy");
				if (i == 7) in[i].X.fp_val.val[1] ^= 4; /* one bad item must not poison the batch */
			}
			int rb = gpu_batch(out, ks, in, NB, rets);
			CHECK(rb == -1, "%s batch: expected overall -1 because of the bad item", names[c]);
			for (int i = 0; i < NB; i++) {
				int r1 = prj_pt_mul(&ref_out[i], &ks[i], &in[i]);
				compare("batch item", names[c], r1, &ref_out[i], rets[i], &out[i]);
			}
		}
		/* ---- ECDSA verify_batch slot */
		if (SECTION('e')) {
			enum { NS = 48 };
			static ec_key_pair kp[NS];
			static u8 sigs[NS][2 * 66], msgs[NS][40];
			const u8 *sp[NS], *mp[NS];
			const ec_pub_key *pk[NS];
			u8 sl[NS];
			u32 ml[NS];
			hash_alg_type ht = (c == 2) ? SHA384 : ((c == 5 || c == 7) ? SHA512 : SHA256);
			for (int i = 0; i < NS; i++) {
				CHECK(!ec_key_pair_gen(&kp[i], &params, ECDSA), "keygen");
				ml[i] = (u32)(1 + (rnd8() % 39));
				for (u32 j = 0; j < ml[i]; j++) msgs[i][j] = rnd8();
				sl[i] = (u8)(2 * qlen);
				CHECK(!ec_sign(sigs[i], sl[i], &kp[i], msgs[i], ml[i], ECDSA, ht, NULL, 0), "ec_sign");
				sp[i] = sigs[i];
				mp[i] = msgs[i];
				pk[i] = &kp[i].pub_key; /* y is a prj_pt_mul output: Z != 1 */
			}
			int r = gpu_vbatch(sp, sl, pk, mp, ml, NS, ECDSA, ht, NULL, NULL, NULL, NULL);
			CHECK(r == 0, "%s verify_batch: valid batch rejected", names[c]);
			sigs[5][3] ^= 0x20;
			msgs[9][0] ^= 1;
			r = gpu_vbatch(sp, sl, pk, mp, ml, NS, ECDSA, ht, NULL, NULL, NULL, NULL);
			CHECK(r == -1, "%s verify_batch: corrupted batch accepted", names[c]);
			signed char v[NS];
			CHECK(gpu_verdicts(v, NS) == NS, "verdict count");
			for (int i = 0; i < NS; i++) {
				int want = ec_verify(sigs[i], sl[i], pk[i], msgs[i], ml[i], ECDSA, ht, NULL, 0) ? -1 : 0;
				CHECK(v[i] == want, "%s verify_batch verdict[%d] = %d, reference ec_verify says %d", names[c], i, v[i], want);
			}
			CHECK(v[5] == -1 && v[9] == -1, "corrupted items not flagged");
			/* ---- ec_verify shim: one kernel launch per ECDSA signature, verdicts equal the reference's; other
			 * schemes are forwarded to the reference's ec_verify (no GPU verification counted) */
			for (int i = 0; i < 12; i++) {
				unsigned long long v0 = gpu_vcount();
				int want = ec_verify(sigs[i], sl[i], pk[i], msgs[i], ml[i], ECDSA, ht, NULL, 0);
				int got = gpu_everify(sigs[i], sl[i], pk[i], msgs[i], ml[i], ECDSA, ht, NULL, 0);
				CHECK(got == want, "%s ec_verify shim item %d: %d vs reference %d", names[c], i, got, want);
				CHECK(gpu_vcount() == v0 + 1, "%s ec_verify shim did not run on the GPU", names[c]);
			}
			CHECK(gpu_everify(sigs[0], (u8)(sl[0] - 1), pk[0], msgs[0], ml[0], ECDSA, ht, NULL, 0) == -1, "short signature accepted");
			{
				/* a call this layer does not serve (ECDSA carrying ancillary data) is forwarded to the reference's
				 * own ec_verify: same verdict, no GPU verification counted */
				unsigned long long v0 = gpu_vcount();
				const u8 ad[3] = { 1, 2, 3 };
				int want = ec_verify(sigs[0], sl[0], pk[0], msgs[0], ml[0], ECDSA, ht, ad, 3);
				CHECK(gpu_everify(sigs[0], sl[0], pk[0], msgs[0], ml[0], ECDSA, ht, ad, 3) == want, "forwarded ec_verify differs");
				CHECK(gpu_vcount() == v0, "a call with ancillary data must be forwarded, not counted as a GPU verification");
			}
			/* ---- a public key that IS the point at infinity: the reference's ec_verify accepts the struct and goes on
			 * with W' = u*G; shim and batch adapter must give the reference's verdict, whatever it is */
			{
				ec_pub_key kinf = kp[0].pub_key;
				CHECK(!prj_pt_zero(&kinf.y), "zero key");
				int want = ec_verify(sigs[0], sl[0], &kinf, msgs[0], ml[0], ECDSA, ht, NULL, 0);
				int got = gpu_everify(sigs[0], sl[0], &kinf, msgs[0], ml[0], ECDSA, ht, NULL, 0);
				CHECK(got == want, "%s key at infinity: shim %d vs reference %d", names[c], got, want);
			}
			/* the generic entry point still reports ECDSA batch as unsupported in the unmodified reference */
			CHECK(ec_verify_batch(sp, sl, pk, mp, ml, NS, ECDSA, ht, NULL, NULL, NULL, NULL) == -1,
			      "reference ec_verify_batch(ECDSA) unexpectedly supported");
			/* ... while the drop-in's ec_verify_batch / is_verify_batch_mode_supported (the reference's prototypes)
			 * serve it, and forward what the layer does not handle (EdDSA) */
			{
				vbatch_fn gpu_generic = (vbatch_fn)dlsym(h, "ec_verify_batch");
				int (*gpu_supported)(ec_alg_type, int *) = (int (*)(ec_alg_type, int *))dlsym(h, "is_verify_batch_mode_supported");
				int chk = -1, refchk = -1;
				CHECK(gpu_generic != NULL && gpu_supported != NULL, "missing ec_verify_batch / is_verify_batch_mode_supported");
				CHECK(gpu_generic(sp, sl, pk, mp, ml, NS, ECDSA, ht, NULL, NULL, NULL, NULL) == -1, "generic batch: corrupted batch accepted");
				sigs[5][3] ^= 0x20;
				msgs[9][0] ^= 1;
				unsigned long long v0 = gpu_vcount();
				CHECK(gpu_generic(sp, sl, pk, mp, ml, NS, ECDSA, ht, NULL, NULL, NULL, NULL) == 0, "generic batch: valid batch rejected");
				CHECK(gpu_vcount() == v0 + NS, "generic ec_verify_batch did not run on the GPU");
				CHECK(gpu_supported(ECDSA, &chk) == 0 && chk == 1, "ECDSA batch mode not reported");
				CHECK(is_verify_batch_mode_supported(ECDSA, &refchk) == 0 && refchk == 0, "reference reports ECDSA batch mode");
				CHECK(gpu_supported(EDDSA25519, &chk) == 0 && is_verify_batch_mode_supported(EDDSA25519, &refchk) == 0 && chk == refchk,
				      "forwarded is_verify_batch_mode_supported(EDDSA25519) differs from the reference");
				CHECK(gpu_generic(sp, sl, pk, mp, ml, NS, EDDSA25519, ht, NULL, NULL, NULL, NULL) ==
				      ec_verify_batch(sp, sl, pk, mp, ml, NS, EDDSA25519, ht, NULL, NULL, NULL, NULL), "forwarded ec_verify_batch(EDDSA25519) differs");
				CHECK(gpu_supported(ECGDSA, &chk) == 0 && chk == 1 && gpu_supported(SM2, &chk) == 0 && chk == 1 &&
				      gpu_supported(BIGN, &chk) == 0 && chk == 1, "batch mode of the double-scalar schemes not reported");
			}
		}
		/* ---- ECFSDSA in the same slot: against the reference's ec_verify, item by item */
		if (SECTION('f')) {
			enum { NF = 24 };
			static ec_key_pair kp[NF];
			static u8 sigs[NF][3 * 66], msgs[NF][40];
			const u8 *sp[NF], *mp[NF];
			const ec_pub_key *pk[NF];
			u8 sl[NF], plen = (u8)BYTECEIL(params.ec_fp.p_bitlen);
			u32 ml[NF];
			hash_alg_type ht = (c == 2) ? SHA384 : ((c == 5 || c == 7) ? SHA512 : SHA256);
			for (int i = 0; i < NF; i++) {
				CHECK(!ec_key_pair_gen(&kp[i], &params, ECFSDSA), "ecfsdsa keygen");
				ml[i] = (u32)(1 + (rnd8() % 39));
				for (u32 j = 0; j < ml[i]; j++) msgs[i][j] = rnd8();
				sl[i] = (u8)(2 * plen + qlen);
				CHECK(!ec_sign(sigs[i], sl[i], &kp[i], msgs[i], ml[i], ECFSDSA, ht, NULL, 0), "ec_sign ECFSDSA");
				sp[i] = sigs[i];
				mp[i] = msgs[i];
				pk[i] = &kp[i].pub_key;
			}
			int r = gpu_fsbatch(sp, sl, pk, mp, ml, NF, ECFSDSA, ht, NULL, NULL, NULL, NULL);
			CHECK(r == 0, "%s ecfsdsa verify_batch: valid batch rejected", names[c]);
			sigs[3][2 * plen + 1] ^= 0x10; /* s */
			sigs[7][0] ^= 1;               /* r off the curve */
			msgs[11][0] ^= 1;
			r = gpu_fsbatch(sp, sl, pk, mp, ml, NF, ECFSDSA, ht, NULL, NULL, NULL, NULL);
			CHECK(r == -1, "%s ecfsdsa verify_batch: corrupted batch accepted", names[c]);
			signed char v[NF];
			CHECK(gpu_verdicts(v, NF) == NF, "verdict count");
			for (int i = 0; i < NF; i++) {
				int want = ec_verify(sigs[i], sl[i], pk[i], msgs[i], ml[i], ECFSDSA, ht, NULL, 0) ? -1 : 0;
				CHECK(v[i] == want, "%s ecfsdsa verdict[%d] = %d, reference ec_verify says %d", names[c], i, v[i], want);
			}
			CHECK(v[3] == -1 && v[7] == -1 && v[11] == -1, "corrupted ECFSDSA items not flagged");
			/* a key at infinity: e*Y = infinity, W' = s*G (sig/ecfsdsa.c:597-600), so (r = k*G, s = k) verifies under it
			 * in the reference whatever the message; the layer must agree, and reject it once s is off by one */
			{
				static ec_pub_key kinf;
				nn k;
				prj_pt R;
				u8 sg[3 * 66], kb[66];
				const u8 *sp1[1] = { sg }, *mp1[1] = { msgs[0] };
				const ec_pub_key *pk1[1] = { &kinf };
				kinf = kp[0].pub_key;
				CHECK(!prj_pt_zero(&kinf.y), "zero key");
				for (int j = 0; j < qlen; j++) kb[j] = rnd8();
				CHECK(!nn_init_from_buf(&k, kb, qlen) && !nn_mod(&k, &k, &params.ec_gen_order), "k");
				CHECK(!prj_pt_mul(&R, &k, &params.ec_gen) && !prj_pt_export_to_aff_buf(&R, sg, (u32)(2 * plen)), "kG");
				CHECK(!nn_export_to_buf(sg + 2 * plen, qlen, &k), "s = k");
				for (int pass = 0; pass < 2; pass++) {
					int want = ec_verify(sg, sl[0], &kinf, msgs[0], ml[0], ECFSDSA, ht, NULL, 0);
					CHECK(want == (pass ? -1 : 0), "%s: reference verdict %d on the crafted ECFSDSA signature, pass %d", names[c], want, pass);
					CHECK(gpu_everify(sg, sl[0], &kinf, msgs[0], ml[0], ECFSDSA, ht, NULL, 0) == want,
					      "%s ECFSDSA key at infinity: shim differs from the reference (%d)", names[c], want);
					CHECK(gpu_fsbatch(sp1, sl, pk1, mp1, ml, 1, ECFSDSA, ht, NULL, NULL, NULL, NULL) == want,
					      "%s ECFSDSA key at infinity: batch adapter differs from the reference (%d)", names[c], want);
					sg[2 * plen + qlen - 1] ^= 1;
				}
			}
		}
		/* ---- ECSDSA / ECOSDSA: batch adapter (W' = sG + eY on the device, hashing of W' with the reference's src/hash)
		 * and ec_verify shim against the reference's ec_verify */
		for (int alt = 0; alt < 3 && SECTION('s'); alt++) {
			enum { NSD = 20 };
			const ec_alg_type alg = alt == 0 ? ECSDSA : (alt == 1 ? ECOSDSA : ECKCDSA);
			static ec_key_pair kp[NSD];
			static u8 sigs[NSD][64 + 66 + 8], msgs[NSD][40];
			const u8 *sp[NSD], *mp[NSD];
			const ec_pub_key *pk[NSD];
			u8 sl[NSD], sgl = 0;
			u32 ml[NSD];
			hash_alg_type ht = (c == 2) ? SHA384 : ((c == 5 || c == 7) ? SHA512 : SHA256);
			vbatch_fn gpu_sdbatch = (vbatch_fn)dlsym(h, alt == 2 ? "eccb200_dropin_eckcdsa_verify_batch"
									      : "eccb200_dropin_ecsdsa_verify_batch");
			CHECK(gpu_sdbatch != NULL, "missing eccb200_dropin_ec(k|s)cdsa_verify_batch");
			if (!gpu_sdbatch) break;
			CHECK(!ec_get_sig_len(&params, alg, ht, &sgl), "siglen");
			for (int i = 0; i < NSD; i++) {
				CHECK(!ec_key_pair_gen(&kp[i], &params, alg), "ecsdsa keygen");
				ml[i] = (u32)(1 + (rnd8() % 39));
				for (u32 j = 0; j < ml[i]; j++) msgs[i][j] = rnd8();
				sl[i] = sgl;
				CHECK(!ec_sign(sigs[i], sl[i], &kp[i], msgs[i], ml[i], alg, ht, NULL, 0), "ec_sign ECSDSA");
				sp[i] = sigs[i];
				mp[i] = msgs[i];
				pk[i] = &kp[i].pub_key;
			}
			int r = gpu_sdbatch(sp, sl, pk, mp, ml, NSD, alg, ht, NULL, NULL, NULL, NULL);
			CHECK(r == 0, "%s scheme %d verify_batch: valid batch rejected", names[c], (int)alg);
			sigs[4][1] ^= 0x08;        /* r */
			sigs[8][sgl - 1] ^= 1;     /* s */
			msgs[12][0] ^= 1;
			r = gpu_sdbatch(sp, sl, pk, mp, ml, NSD, alg, ht, NULL, NULL, NULL, NULL);
			CHECK(r == -1, "%s scheme %d verify_batch: corrupted batch accepted", names[c], (int)alg);
			signed char v[NSD];
			CHECK(gpu_verdicts(v, NSD) == NSD, "verdict count");
			for (int i = 0; i < NSD; i++) {
				int want = ec_verify(sigs[i], sl[i], pk[i], msgs[i], ml[i], alg, ht, NULL, 0) ? -1 : 0;
				CHECK(v[i] == want, "%s scheme %d verdict[%d] = %d, reference ec_verify says %d", names[c], (int)alg, i, v[i], want);
				int got = gpu_everify(sigs[i], sl[i], pk[i], msgs[i], ml[i], alg, ht, NULL, 0) ? -1 : 0;
				CHECK(got == want, "%s scheme %d ec_verify shim item %d: %d vs %d", names[c], (int)alg, i, got, want);
			}
			CHECK(v[4] == -1 && v[8] == -1 && v[12] == -1, "corrupted items of scheme %d not flagged", (int)alg);
		}
		/* ---- ECGDSA / ECRDSA / SM2 / BIGN / DBIGN: W' = a*G + b*Y on the device, mod-q scalar preparation (one shared
		 * inversion per chunk), hashes and comparison on the host; batch adapters, the generic ec_verify_batch and the
		 * ec_verify shim against the reference's ec_verify, with the ancillary data SM2 (signer ID) and BIGN (hash OID) need */
		for (int alt = 0; alt < 5 && SECTION('d'); alt++) {
			enum { NG = 20 };
			const ec_alg_type algs[5] = { ECGDSA, ECRDSA, SM2, BIGN, DBIGN };
			const char *fns[5] = { "eccb200_dropin_ecgdsa_verify_batch", "eccb200_dropin_ecrdsa_verify_batch",
					       "eccb200_dropin_sm2_verify_batch", "eccb200_dropin_bign_verify_batch",
					       "eccb200_dropin_bign_verify_batch" };
			const ec_alg_type alg = algs[alt];
			static ec_key_pair kp[NG];
			static u8 sigs[NG][2 * 66 + 8], msgs[NG][40];
			const u8 *sp[NG], *mp[NG], *ap[NG];
			const ec_pub_key *pk[NG];
			u8 sl[NG], sgl = 0;
			u32 ml[NG];
			u16 al[NG];
			/* SM2: the signer's ID; BIGN: oid_len || t_len || oid || t (sig/bign_common.c:97-140) */
			static const u8 sm2_id[] = "libecc-b200@example";
			static const u8 bign_ad[] = { 0x00, 0x0b, 0x00, 0x04, 0x06, 0x09, 0x2a, 0x70, 0x00, 0x02, 0x00, 0x22, 0x65, 0x1f, 0x51,
						      0xde, 0xad, 0xbe, 0xef };
			const u8 *ad = alg == SM2 ? sm2_id : ((alg == BIGN || alg == DBIGN) ? bign_ad : NULL);
			const u16 adl = alg == SM2 ? (u16)(sizeof(sm2_id) - 1) : ((alg == BIGN || alg == DBIGN) ? (u16)sizeof(bign_ad) : 0);
			hash_alg_type ht = (c == 2) ? SHA384 : ((c == 5 || c == 7) ? SHA512 : SHA256);
			if (alg == SM2 && (c & 1)) ht = SM3;
			if (alg == ECRDSA && (c & 1)) ht = (c == 5 || c == 7) ? STREEBOG512 : STREEBOG256;
			vbatch_fn gpu_dsbatch = (vbatch_fn)dlsym(h, fns[alt]);
			vbatch_fn gpu_generic = (vbatch_fn)dlsym(h, "ec_verify_batch");
			CHECK(gpu_dsbatch != NULL && gpu_generic != NULL, "missing %s", fns[alt]);
			if (!gpu_dsbatch || !gpu_generic) break;
			CHECK(!ec_get_sig_len(&params, alg, ht, &sgl), "siglen");
			int usable = 1;
			for (int i = 0; i < NG && usable; i++) {
				if (ec_key_pair_gen(&kp[i], &params, alg)) { usable = 0; break; }
				ml[i] = (u32)(1 + (rnd8() % 39));
				for (u32 j = 0; j < ml[i]; j++) msgs[i][j] = rnd8();
				sl[i] = sgl;
				if (ec_sign(sigs[i], sl[i], &kp[i], msgs[i], ml[i], alg, ht, ad, adl)) { usable = 0; break; }
				sp[i] = sigs[i];
				mp[i] = msgs[i];
				pk[i] = &kp[i].pub_key;
				ap[i] = ad;
				al[i] = adl;
			}
			if (!usable) {
				printf("note: the reference does not sign scheme %d on %s here; skipped\n", (int)alg, names[c]);
				continue;
			}
			unsigned long long v0 = gpu_vcount();
			int r = gpu_dsbatch(sp, sl, pk, mp, ml, NG, alg, ht, ad ? ap : NULL, ad ? al : NULL, NULL, NULL);
			CHECK(r == 0, "%s scheme %d verify_batch: valid batch rejected", names[c], (int)alg);
			CHECK(gpu_vcount() == v0 + NG, "%s scheme %d verify_batch did not run on the GPU", names[c], (int)alg);
			r = gpu_generic(sp, sl, pk, mp, ml, NG, alg, ht, ad ? ap : NULL, ad ? al : NULL, NULL, NULL);
			CHECK(r == 0, "%s scheme %d generic ec_verify_batch: valid batch rejected", names[c], (int)alg);
			sigs[4][1] ^= 0x08;              /* first half: r (s0 for BIGN) */
			sigs[8][sgl - 2] ^= 1;           /* second half: s (s1) */
			msgs[12][0] ^= 1;
			{ /* second half := q - first half (SM2: r + s = q, t = 0, sig/sm2.c:660; elsewhere just another forgery) */
				nn a, d;
				if (!nn_init_from_buf(&a, sigs[10], qlen) && !nn_mod(&a, &a, &params.ec_gen_order) &&
				    !nn_sub(&d, &params.ec_gen_order, &a) && sgl >= 2 * qlen) {
					CHECK(!nn_export_to_buf(sigs[10], qlen, &a) && !nn_export_to_buf(sigs[10] + qlen, qlen, &d), "r + s = q");
				}
			}
			memset(sigs[14], 0xff, sgl);     /* out of range */
			memset(sigs[15], 0, sgl);        /* zero */
			static ec_pub_key kinf, koff;    /* a key at infinity, a key off the curve */
			kinf = kp[16].pub_key;
			CHECK(!prj_pt_zero(&kinf.y), "zero key");
			pk[16] = &kinf;
			koff = kp[17].pub_key;
			koff.y.X.fp_val.val[0] ^= 2;
			pk[17] = &koff;
			static const u8 other_id[] = "someone else";
			static const u8 short_ad[] = { 0x00, 0x09, 0x00, 0x01, 0x06 };
			if (ad) {
				ap[18] = alg == SM2 ? other_id : short_ad; /* another signer ID / an OID record that overruns */
				al[18] = alg == SM2 ? (u16)(sizeof(other_id) - 1) : (u16)sizeof(short_ad);
				ap[19] = NULL;                             /* no ancillary data at all */
				al[19] = 0;
			}
			r = gpu_dsbatch(sp, sl, pk, mp, ml, NG, alg, ht, ad ? ap : NULL, ad ? al : NULL, NULL, NULL);
			CHECK(r == -1, "%s scheme %d verify_batch: corrupted batch accepted", names[c], (int)alg);
			signed char v[NG];
			CHECK(gpu_verdicts(v, NG) == NG, "verdict count");
			for (int i = 0; i < NG; i++) {
				const u8 *adi = ad ? ap[i] : NULL;
				const u16 adli = ad ? al[i] : 0;
				int want = ec_verify(sigs[i], sl[i], pk[i], msgs[i], ml[i], alg, ht, adi, adli) ? -1 : 0;
				CHECK(v[i] == want, "%s scheme %d verdict[%d] = %d, reference ec_verify says %d", names[c], (int)alg, i, v[i], want);
				unsigned long long v1 = gpu_vcount();
				int got = gpu_everify(sigs[i], sl[i], pk[i], msgs[i], ml[i], alg, ht, adi, adli) ? -1 : 0;
				CHECK(got == want, "%s scheme %d ec_verify shim item %d: %d vs %d", names[c], (int)alg, i, got, want);
				if (i < 4) CHECK(gpu_vcount() == v1 + 1, "%s scheme %d ec_verify shim did not run on the GPU", names[c], (int)alg);
			}
			CHECK(v[4] == -1 && v[8] == -1 && v[12] == -1 && v[14] == -1 && v[15] == -1 && v[17] == -1 &&
			      (sgl < 2 * qlen || v[10] == -1), "corrupted items of scheme %d not flagged", (int)alg);
			CHECK(v[0] == 0 && v[1] == 0 && v[ad ? 13 : 19] == 0, "valid items of scheme %d rejected", (int)alg);
		}
		/* ---- BIP0340 in the same slot and through the ec_verify shim: against the reference's ec_verify */
		if (SECTION('b')) {
			enum { NB3 = 20 };
			static ec_key_pair kp[NB3];
			static u8 sigs[NB3][2 * 66], msgs[NB3][40];
			const u8 *sp[NB3], *mp[NB3];
			const ec_pub_key *pk[NB3];
			u8 sl[NB3], plen = (u8)BYTECEIL(params.ec_fp.p_bitlen);
			u32 ml[NB3];
			hash_alg_type ht = (c == 2) ? SHA384 : ((c == 5 || c == 7) ? SHA512 : SHA256);
			vbatch_fn gpu_bipbatch = (vbatch_fn)dlsym(h, "eccb200_dropin_bip0340_verify_batch");
			CHECK(gpu_bipbatch != NULL, "missing eccb200_dropin_bip0340_verify_batch");
			int usable = gpu_bipbatch != NULL;
			for (int i = 0; i < NB3 && usable; i++) {
				if (ec_key_pair_gen(&kp[i], &params, BIP0340)) { usable = 0; break; }
				ml[i] = (u32)(1 + (rnd8() % 39));
				for (u32 j = 0; j < ml[i]; j++) msgs[i][j] = rnd8();
				sl[i] = (u8)(plen + qlen);
				if (ec_sign(sigs[i], sl[i], &kp[i], msgs[i], ml[i], BIP0340, ht, NULL, 0)) { usable = 0; break; }
				sp[i] = sigs[i];
				mp[i] = msgs[i];
				pk[i] = &kp[i].pub_key;
			}
			if (usable) {
				int r = gpu_bipbatch(sp, sl, pk, mp, ml, NB3, BIP0340, ht, NULL, NULL, NULL, NULL);
				CHECK(r == 0, "%s bip0340 verify_batch: valid batch rejected", names[c]);
				sigs[2][plen + 1] ^= 0x10; /* s */
				sigs[6][3] ^= 1;           /* r */
				msgs[9][0] ^= 1;
				r = gpu_bipbatch(sp, sl, pk, mp, ml, NB3, BIP0340, ht, NULL, NULL, NULL, NULL);
				CHECK(r == -1, "%s bip0340 verify_batch: corrupted batch accepted", names[c]);
				signed char v[NB3];
				CHECK(gpu_verdicts(v, NB3) == NB3, "verdict count");
				for (int i = 0; i < NB3; i++) {
					int want = ec_verify(sigs[i], sl[i], pk[i], msgs[i], ml[i], BIP0340, ht, NULL, 0) ? -1 : 0;
					CHECK(v[i] == want, "%s bip0340 verdict[%d] = %d, reference ec_verify says %d", names[c], i, v[i], want);
					int got = gpu_everify(sigs[i], sl[i], pk[i], msgs[i], ml[i], BIP0340, ht, NULL, 0) ? -1 : 0;
					CHECK(got == want, "%s bip0340 ec_verify shim item %d: %d vs %d", names[c], i, got, want);
				}
				CHECK(v[2] == -1 && v[6] == -1 && v[9] == -1, "corrupted BIP0340 items not flagged");
			} else {
				printf("note: the reference does not sign BIP0340 on %s here; skipped\n", names[c]);
			}
		}
		printf("direct %s done, failures so far %d\n", names[c], failures);
	}
	return failures != 0;
}


/* kats mode: every fixed-vector test case of the reference's self tests (src/tests/ec_self_tests_core.h:4915-, the table
 * `ec_self_tests vectors` walks) through the drop-in's ec_verify: the expected signature must verify, a flipped message
 * bit must not, both exactly as the reference's own ec_verify says; schemes / curves the layer does not serve (EdDSA,
 * GOST and other curves) are forwarded and still agree.  Prints how many cases ran on the GPU per scheme. */
static int run_kats(const char *dropin_path)
{
	void *h = dlopen(dropin_path, RTLD_NOW | RTLD_LOCAL);
	if (!h) {
		printf("FAIL dlopen %s: %s\n", dropin_path, dlerror());
		return 1;
	}
	everify_fn gpu_everify = (everify_fn)dlsym(h, "eccb200_dropin_ec_verify");
	count_fn gpu_vcount = (count_fn)dlsym(h, "eccb200_dropin_verify_count");
	vbatch_fn gpu_generic = (vbatch_fn)dlsym(h, "ec_verify_batch");
	if (!gpu_everify || !gpu_vcount || !gpu_generic) {
		printf("FAIL missing drop-in symbols\n");
		return 1;
	}
	const unsigned ncases = (unsigned)(sizeof(ec_fixed_vector_tests) / sizeof(ec_fixed_vector_tests[0]));
	unsigned served[32] = { 0 }, total[32] = { 0 }, ran = 0;
	for (unsigned t = 0; t < ncases; t++) {
		const ec_test_case *c = ec_fixed_vector_tests[t];
		ec_params params;
		ec_key_pair kp;
		static u8 msg[4096];
		/* HARNESS_KATS_THIN (the CPU run of the tests): every third of the 47 ECDSA / DECDSA cases, all of the others */
		if (getenv("HARNESS_KATS_THIN") && (c->sig_type == ECDSA || c->sig_type == DECDSA) && (t % 3)) continue;
		if (c->msglen > sizeof(msg) || c->msglen == 0) continue;
		if (import_params(&params, c->ec_str_p)) continue;
		const int eddsa = c->sig_type == EDDSA25519 || c->sig_type == EDDSA25519CTX || c->sig_type == EDDSA25519PH ||
				  c->sig_type == EDDSA448 || c->sig_type == EDDSA448PH;
		/* key import as the reference's own self test does it (src/tests/ec_self_tests_core.c:752-783) */
		if (eddsa ? eddsa_import_key_pair_from_priv_key_buf(&kp, c->priv_key, c->priv_key_len, &params, c->sig_type)
			  : ec_key_pair_import_from_priv_key_buf(&kp, &params, c->priv_key, c->priv_key_len, c->sig_type))
			continue;
		memcpy(msg, c->msg, c->msglen);
		const unsigned long long v0 = gpu_vcount();
		int want = ec_verify(c->exp_sig, c->exp_siglen, &kp.pub_key, msg, c->msglen, c->sig_type, c->hash_type, c->adata, c->adata_len);
		int got = gpu_everify(c->exp_sig, c->exp_siglen, &kp.pub_key, msg, c->msglen, c->sig_type, c->hash_type, c->adata, c->adata_len);
		CHECK(want == 0, "%s: the reference rejects its own expected signature (%d)", c->name, want);
		CHECK(got == want, "%s: drop-in ec_verify %d, reference %d", c->name, got, want);
		msg[c->msglen / 2] ^= 0x04;
		want = ec_verify(c->exp_sig, c->exp_siglen, &kp.pub_key, msg, c->msglen, c->sig_type, c->hash_type, c->adata, c->adata_len);
		got = gpu_everify(c->exp_sig, c->exp_siglen, &kp.pub_key, msg, c->msglen, c->sig_type, c->hash_type, c->adata, c->adata_len);
		CHECK(want == -1 && got == want, "%s (message altered): drop-in %d, reference %d", c->name, got, want);
		msg[c->msglen / 2] ^= 0x04;
		/* the same case as a batch of one through the generic ec_verify_batch (adata as arrays) */
		{
			const u8 *sp[1] = { c->exp_sig }, *mp[1] = { msg }, *ap[1] = { c->adata };
			const ec_pub_key *pk[1] = { &kp.pub_key };
			u8 sl[1] = { c->exp_siglen };
			u32 ml[1] = { c->msglen };
			u16 al[1] = { c->adata_len };
			int chk = 0;
			int (*gpu_supported)(ec_alg_type, int *) = (int (*)(ec_alg_type, int *))dlsym(h, "is_verify_batch_mode_supported");
			/* (the first two cases of every scheme: the batch entry shares the scheme code with the shim) */
			if ((unsigned)c->sig_type < 32 && total[c->sig_type] < 2 && gpu_supported &&
			    !gpu_supported(c->sig_type, &chk) && chk && gpu_vcount() > v0) {
				static verify_batch_scratch_pad pad[16];
				u32 padlen = sizeof(pad);
				int b = gpu_generic(sp, sl, pk, mp, ml, 1, c->sig_type, c->hash_type, c->adata ? ap : NULL, c->adata ? al : NULL, pad, &padlen);
				CHECK(b == 0, "%s: generic ec_verify_batch of one rejects the expected signature (%d)", c->name, b);
			}
		}
		if ((unsigned)c->sig_type < 32) {
			total[c->sig_type]++;
			if (gpu_vcount() > v0) served[c->sig_type]++;
		}
		ran++;
	}
	printf("kats: %u of %u fixed-vector cases of the reference's self tests ran; served by the GPU per ec_alg_type:", ran, ncases);
	for (int a = 0; a < 32; a++)
		if (total[a]) printf(" %d:%u/%u", a, served[a], total[a]);
	printf("\n");
	CHECK(ran >= (getenv("HARNESS_KATS_THIN") ? 90u : 120u), "too few known-answer cases ran (%u)", ran);
	return failures != 0;
}


/* fuzz mode (host logic, differential): reference-made signatures of every scheme the layer serves are mutated - single
 * bit flips, halves forced to 0 / q - 1 / q / all-ones, lengths off by one, altered or missing ancillary data, keys at
 * infinity / off the curve / of another scheme / with a broken magic - and every mutant goes through the reference's
 * ec_verify and the drop-in's: the two verdicts must be equal.  Deterministic (seeded); `iters` mutants per scheme. */
static int run_fuzz(const char *dropin_path, const char *curve, u32 iters, const char *hash_name)
{
	const hash_mapping *fuzz_hm = NULL;
	if (get_hash_by_name(hash_name, &fuzz_hm) || !fuzz_hm) {
		printf("FAIL unknown hash %s\n", hash_name);
		return 1;
	}
	const hash_alg_type HT = fuzz_hm->type; /* digests shorter and longer than the order exercise the truncation rules */
	void *h = dlopen(dropin_path, RTLD_NOW | RTLD_LOCAL);
	if (!h) return 1;
	everify_fn gpu_everify = (everify_fn)dlsym(h, "eccb200_dropin_ec_verify");
	count_fn gpu_vcount = (count_fn)dlsym(h, "eccb200_dropin_verify_count");
	vbatch_fn gpu_generic = (vbatch_fn)dlsym(h, "ec_verify_batch");
	verdicts_fn gpu_verdicts = (verdicts_fn)dlsym(h, "eccb200_dropin_last_verdicts");
	if (!gpu_everify || !gpu_vcount || !gpu_generic || !gpu_verdicts) return 1;
	/* the same mutants, sixteen at a time, also go through ec_verify_batch: per-item verdicts and the 0 / -1 of the call */
	enum { FB = 16 };
	static u8 b_sig[FB][3 * 66 + 2], b_msg[FB][40], b_ad[FB][32];
	static ec_pub_key b_pk[FB];
	const u8 *b_sp[FB], *b_mp[FB], *b_ap[FB];
	const ec_pub_key *b_pkp[FB];
	u8 b_sl[FB];
	u32 b_ml[FB];
	u16 b_al[FB];
	int b_want[FB];
	unsigned long long batches = 0;
	ec_params params;
	CHECK(!load_params(&params, curve), "params");
	for (const char *c = curve; *c; c++) rng_state = rng_state * 31 + (unsigned char)*c; /* another sequence per curve */
	const u8 qlen = (u8)BYTECEIL(params.ec_gen_order_bitlen);
	u8 qb[66], qm1[66];
	nn t;
	CHECK(!nn_export_to_buf(qb, qlen, &params.ec_gen_order), "q");
	CHECK(!nn_init(&t, 0) && !nn_dec(&t, &params.ec_gen_order) && !nn_export_to_buf(qm1, qlen, &t), "q - 1");
	static const ec_alg_type algs[12] = { ECDSA, DECDSA, ECFSDSA, BIP0340, ECKCDSA, ECSDSA, ECOSDSA, ECGDSA, ECRDSA, SM2, BIGN, DBIGN };
	static const u8 sm2_id[] = "fuzz@libecc-b200";
	static const u8 bign_ad[] = { 0x00, 0x0b, 0x00, 0x04, 0x06, 0x09, 0x2a, 0x70, 0x00, 0x02, 0x00, 0x22, 0x65, 0x1f, 0x51,
				      0xde, 0xad, 0xbe, 0xef };
	unsigned long long accepted = 0, total = 0, on_gpu = 0;
	for (int a = 0; a < 12; a++) {
		const ec_alg_type alg = algs[a];
		const int needs_ad = alg == SM2 || alg == BIGN || alg == DBIGN;
		const u8 *ad = alg == SM2 ? sm2_id : (needs_ad ? bign_ad : NULL);
		const u16 adl = alg == SM2 ? (u16)(sizeof(sm2_id) - 1) : (needs_ad ? (u16)sizeof(bign_ad) : 0);
		enum { NK = 3 };
		static ec_key_pair kp[NK];
		static u8 sig0[NK][3 * 66], msg0[NK][40];
		u32 ml0[NK];
		u8 sgl = 0;
		int usable = !ec_get_sig_len(&params, alg, HT, &sgl) && sgl <= 3 * 66 - 2;
		for (int i = 0; i < NK && usable; i++) {
			ml0[i] = (u32)(1 + (rnd8() % 39));
			for (u32 j = 0; j < ml0[i]; j++) msg0[i][j] = rnd8();
			usable = !ec_key_pair_gen(&kp[i], &params, alg) && !ec_sign(sig0[i], sgl, &kp[i], msg0[i], ml0[i], alg, HT, ad, adl);
		}
		if (!usable) {
			printf("note: scheme %d not usable on %s here; skipped\n", (int)alg, curve);
			continue;
		}
		int nb = 0;
		for (u32 it = 0; it < iters; it++) {
			const int k = (int)(rnd8() % NK);
			u8 sig[3 * 66 + 2], msg[40], adbuf[32];
			static ec_pub_key pk;
			const u8 *adp = ad;
			u16 adlen = adl;
			u8 sl = sgl;
			u32 ml = ml0[k];
			memset(sig, 0, sizeof(sig));
			memcpy(sig, sig0[k], sgl);
			memcpy(msg, msg0[k], ml);
			pk = kp[k].pub_key;
			const unsigned half = sgl / 2u, kind = rnd8() % 16u, which = rnd8() & 1u;
			const unsigned off = which ? (unsigned)(sgl - qlen) : 0u; /* the scalar-sized field at either end */
			switch (kind) {
			case 0: break;                                                    /* untouched: must still verify */
			case 1: sig[rnd8() % sgl] ^= (u8)(1u << (rnd8() & 7)); break;     /* one bit of the signature */
			case 2: msg[rnd8() % ml] ^= (u8)(1u << (rnd8() & 7)); break;      /* one bit of the message */
			case 3: memset(sig + off, 0, qlen); break;                        /* field := 0 */
			case 4: memcpy(sig + off, qb, qlen); break;                       /* field := q */
			case 5: memcpy(sig + off, qm1, qlen); break;                      /* field := q - 1 */
			case 6: memset(sig + off, 0xff, qlen); break;                     /* field := 2^(8 qlen) - 1 */
			case 7: sl = (u8)(sgl - 1); break;                                /* one byte short */
			case 8: sl = (u8)(sgl + 1); break;                                /* one byte long */
			case 9: memset(sig, 0, half); break;                              /* leading half zero */
			case 10:                                                          /* ancillary data altered / shortened / absent */
				if (needs_ad) {
					const unsigned m3 = rnd8() % 3u;
					memcpy(adbuf, ad, adl);
					if (m3 == 0) { adbuf[rnd8() % adl] ^= 1; adp = adbuf; }
					else if (m3 == 1) { adp = adbuf; adlen = (u16)(rnd8() % adl); }
					else { adp = NULL; adlen = 0; }
				} else if (rnd8() & 1) { /* a scheme that takes none is handed some: forwarded, must still agree */
					adbuf[0] = 1;
					adp = adbuf;
					adlen = 1;
				}
				break;
			case 11: CHECK(!prj_pt_zero(&pk.y), "zero key"); break;           /* key at infinity */
			case 12: pk.y.X.fp_val.val[0] ^= 4; break;                        /* key off the curve */
			case 13: pk.key_type = algs[(a + 1) % 12]; break;                 /* key of another scheme */
			case 14: pk.magic ^= 1; break;                                    /* not an initialised key */
			default: ml = 0; break;                                           /* empty message */
			}
			const unsigned long long v0 = gpu_vcount();
			const int want = ec_verify(sig, sl, &pk, msg, ml, alg, HT, adp, adlen);
			const int got = gpu_everify(sig, sl, &pk, msg, ml, alg, HT, adp, adlen);
			CHECK(want == got, "%s scheme %d mutation %u (iteration %u): drop-in %d, reference %d", curve, (int)alg, kind, it, got, want);
			if (kind == 0) CHECK(want == 0, "%s scheme %d: untouched signature rejected", curve, (int)alg);
			total++;
			accepted += want == 0;
			on_gpu += gpu_vcount() > v0;
			if (!(adp && !needs_ad)) { /* (a batch with ancillary data on a scheme that takes none is forwarded whole) */
				memcpy(b_sig[nb], sig, sizeof(sig));
				memcpy(b_msg[nb], msg, sizeof(msg));
				if (adp) memcpy(b_ad[nb], adp, adlen);
				b_pk[nb] = pk;
				b_sp[nb] = b_sig[nb];
				b_mp[nb] = b_msg[nb];
				b_ap[nb] = adp ? b_ad[nb] : NULL;
				b_pkp[nb] = &b_pk[nb];
				b_sl[nb] = sl;
				b_ml[nb] = ml;
				b_al[nb] = adlen;
				b_want[nb] = want;
				nb++;
			}
			if (nb == FB) {
				const unsigned long long b0 = gpu_vcount();
				const int rb = gpu_generic(b_sp, b_sl, b_pkp, b_mp, b_ml, FB, alg, HT, needs_ad ? b_ap : NULL,
							   needs_ad ? b_al : NULL, NULL, NULL);
				if (gpu_vcount() == b0 + FB) { /* judged by the engine path (a batch without one usable key is forwarded) */
					signed char v[FB];
					int all = 0;
					CHECK(gpu_verdicts(v, FB) == FB, "verdict count");
					for (int i = 0; i < FB; i++) {
						CHECK(v[i] == b_want[i], "%s scheme %d batch item %d: %d, reference ec_verify %d", curve, (int)alg, i, v[i], b_want[i]);
						if (b_want[i]) all = -1;
					}
					CHECK(rb == all, "%s scheme %d: ec_verify_batch returned %d, items say %d", curve, (int)alg, rb, all);
					batches++;
				}
				nb = 0;
			}
		}
	}
	printf("fuzz %s %s: %llu mutants, %llu accepted by both, %llu judged by the engine path, the rest forwarded; %llu batches of %d through ec_verify_batch\n",
	       curve, hash_name, total, accepted, on_gpu, batches, (int)FB);
	return failures != 0;
}


/* fuzzmul mode: prj_pt_mul on random scalars of every nn width (0 to 27 words, top words set) and on reference-made
 * points in projective form (Z != 1), at infinity, off the curve, with out == in - return code, infinity flag and
 * affine coordinates must equal the reference's. */
static int run_fuzzmul(const char *dropin_path, const char *curve, u32 iters)
{
	void *h = dlopen(dropin_path, RTLD_NOW | RTLD_LOCAL);
	if (!h) return 1;
	mul_fn gpu_mul = (mul_fn)dlsym(h, "prj_pt_mul");
	if (!gpu_mul) return 1;
	ec_params params;
	CHECK(!load_params(&params, curve), "params");
	for (const char *c = curve; *c; c++) rng_state = rng_state * 131 + (unsigned char)*c;
	prj_pt base;
	CHECK(!prj_pt_copy(&base, &params.ec_gen), "copy");
	for (u32 it = 0; it < iters; it++) {
		u8 kb[216];
		const u16 kl = (u16)(rnd8() % 5 == 0 ? 216 - (rnd8() % 9) : (rnd8() % 80));
		for (u16 i = 0; i < kl; i++) kb[i] = rnd8();
		if (kl && (rnd8() & 3) == 0) kb[0] = 0xff; /* the top word really used */
		nn k;
		prj_pt in, o_ref, o_gpu;
		CHECK(!nn_init_from_buf(&k, kb, kl), "nn_init_from_buf(%u)", kl);
		CHECK(!prj_pt_copy(&in, &base), "copy");
		const unsigned kind = rnd8() % 12u;
		if (kind == 0) CHECK(!prj_pt_zero(&in), "zero");
		else if (kind == 1) in.Y.fp_val.val[0] ^= 8;           /* off the curve */
		else if (kind == 2) CHECK(!prj_pt_copy(&in, &params.ec_gen), "G"); /* the fixed-base path */
		const int r1 = prj_pt_mul(&o_ref, &k, &in);
		int r2;
		if (kind == 3) { /* out == in */
			prj_pt alias;
			CHECK(!prj_pt_copy(&alias, &in), "copy");
			r2 = gpu_mul(&alias, &k, &alias);
			o_gpu = alias;
		} else {
			r2 = gpu_mul(&o_gpu, &k, &in);
		}
		compare("fuzzmul", curve, r1, &o_ref, r2, &o_gpu);
		if (!r1 && (rnd8() & 1)) { /* walk on: the next base is this (projective, blinded) result, unless it is infinity */
			int z = 0;
			if (!prj_pt_iszero(&o_ref, &z) && !z) base = o_ref;
		}
	}
	printf("fuzzmul %s: %u multiplications agree with the reference\n", curve, iters);
	return failures != 0;
}

/* threads mode: the reference's functions are re-entrant; several host threads call the drop-in at once */
typedef struct {
	mul_fn mul;
	const ec_params *params;
	int id, bad;
} thr_arg;

static void *thr_main(void *p)
{
	thr_arg *a = (thr_arg *)p;
	unsigned long long st = 0x1234567ULL * (unsigned long long)(a->id + 1);
	for (int t = 0; t < 40; t++) {
		u8 kb[66];
		u8 ql = (u8)BYTECEIL(a->params->ec_gen_order_bitlen);
		for (int j = 0; j < ql; j++) {
			st = st * 6364136223846793005ULL + 1442695040888963407ULL;
			kb[j] = (u8)(st >> 56);
		}
		nn k;
		prj_pt o_ref, o_gpu;
		int cmp = 1;
		if (nn_init_from_buf(&k, kb, ql) || prj_pt_mul(&o_ref, &k, &a->params->ec_gen) || a->mul(&o_gpu, &k, &a->params->ec_gen) ||
		    prj_pt_cmp(&o_ref, &o_gpu, &cmp) || cmp)
			a->bad++;
	}
	return NULL;
}

static int run_threads(const char *dropin_path)
{
	void *h = dlopen(dropin_path, RTLD_NOW | RTLD_LOCAL);
	if (!h) return 1;
	mul_fn gpu_mul = (mul_fn)dlsym(h, "prj_pt_mul");
	ec_params p1, p2;
	CHECK(!load_params(&p1, "SECP256R1") && !load_params(&p2, "BRAINPOOLP256R1"), "params");
	enum { NT = 8 };
	pthread_t th[NT];
	thr_arg args[NT];
	for (int i = 0; i < NT; i++) {
		args[i].mul = gpu_mul;
		args[i].params = (i & 1) ? &p2 : &p1;
		args[i].id = i;
		args[i].bad = 0;
		pthread_create(&th[i], NULL, thr_main, &args[i]);
	}
	for (int i = 0; i < NT; i++) {
		pthread_join(th[i], NULL);
		CHECK(args[i].bad == 0, "thread %d: %d wrong results", i, args[i].bad);
	}
	printf("threads: %d threads x 40 concurrent prj_pt_mul calls on two curves agree with the reference\n", NT);
	return failures != 0;
}

/* bench mode: throughput of the reference-facing batch adapter on REAL ec_pub_key structs (one struct per item) */
static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int run_bench(const char *dropin_path, const char *curve, u32 n, const char *scheme, u32 bad_every)
{
	void *h = dlopen(dropin_path, RTLD_NOW | RTLD_LOCAL);
	if (!h) return 1;
	static const struct { const char *name; ec_alg_type alg; const char *fn; } schemes[] = {
		{ "ECDSA", ECDSA, "eccb200_dropin_ecdsa_verify_batch" },     { "ECFSDSA", ECFSDSA, "eccb200_dropin_ecfsdsa_verify_batch" },
		{ "BIP0340", BIP0340, "eccb200_dropin_bip0340_verify_batch" }, { "ECSDSA", ECSDSA, "eccb200_dropin_ecsdsa_verify_batch" },
		{ "ECOSDSA", ECOSDSA, "eccb200_dropin_ecsdsa_verify_batch" }, { "ECKCDSA", ECKCDSA, "eccb200_dropin_eckcdsa_verify_batch" },
		{ "ECGDSA", ECGDSA, "eccb200_dropin_ecgdsa_verify_batch" },   { "ECRDSA", ECRDSA, "eccb200_dropin_ecrdsa_verify_batch" },
		{ "SM2", SM2, "eccb200_dropin_sm2_verify_batch" },            { "BIGN", BIGN, "eccb200_dropin_bign_verify_batch" },
	};
	ec_alg_type alg = ECDSA;
	const char *fn = schemes[0].fn;
	for (unsigned k = 0; k < sizeof(schemes) / sizeof(schemes[0]); k++)
		if (!strcmp(scheme, schemes[k].name)) {
			alg = schemes[k].alg;
			fn = schemes[k].fn;
		}
	/* ancillary data: the signer ID of SM2, the hash OID record of BIGN (the same for every item) */
	static const u8 sm2_id[] = "libecc-b200@example";
	static const u8 bign_ad[] = { 0x00, 0x0b, 0x00, 0x04, 0x06, 0x09, 0x2a, 0x70, 0x00, 0x02, 0x00, 0x22, 0x65, 0x1f, 0x51,
				      0xde, 0xad, 0xbe, 0xef };
	const u8 *ad = alg == SM2 ? sm2_id : (alg == BIGN ? bign_ad : NULL);
	const u16 adl = alg == SM2 ? (u16)(sizeof(sm2_id) - 1) : (alg == BIGN ? (u16)sizeof(bign_ad) : 0);
	vbatch_fn gpu_vbatch = (vbatch_fn)dlsym(h, fn);
	verdicts_fn gpu_verdicts = (verdicts_fn)dlsym(h, "eccb200_dropin_last_verdicts");
	count_fn msm_batches = (count_fn)dlsym(h, "eccb200_dropin_msm_batches");
	CHECK(gpu_vbatch && gpu_verdicts && msm_batches, "missing drop-in symbols");
	ec_params params;
	CHECK(!load_params(&params, curve), "params");
	u8 siglen = 0;
	CHECK(!ec_get_sig_len(&params, alg, SHA256, &siglen), "siglen");
	enum { POOL_MAX = 2048, ML = 32 };
	static ec_key_pair kp[POOL_MAX];
	static u8 psig[POOL_MAX][3 * 66], pmsg[POOL_MAX][ML];
	/* HARNESS_POOL / HARNESS_REPS shrink the signature pool and the repetitions (the CPU run of the tests) */
	const int POOL = getenv("HARNESS_POOL") ? (int)strtoul(getenv("HARNESS_POOL"), NULL, 10) % (POOL_MAX + 1) : POOL_MAX;
	const int reps = getenv("HARNESS_REPS") ? (int)strtoul(getenv("HARNESS_REPS"), NULL, 10) : 4;
	if (POOL < 14 || reps < 2) return 1;
	for (int i = 0; i < POOL; i++) { /* signatures made by the reference */
		CHECK(!ec_key_pair_gen(&kp[i], &params, alg), "keygen");
		for (int j = 0; j < ML; j++) pmsg[i][j] = rnd8();
		CHECK(!ec_sign(psig[i], siglen, &kp[i], pmsg[i], ML, alg, SHA256, ad, adl), "sign");
	}
	/* n independent items: every item owns its ec_pub_key struct, signature and message bytes (tiled from the pool) */
	ec_pub_key *keys = (ec_pub_key *)malloc((size_t)n * sizeof(ec_pub_key));
	u8 *sigs = (u8 *)malloc((size_t)n * siglen), *msgs = (u8 *)malloc((size_t)n * ML);
	const u8 **sp = malloc((size_t)n * sizeof(*sp)), **mp = malloc((size_t)n * sizeof(*mp));
	const ec_pub_key **pk = malloc((size_t)n * sizeof(*pk));
	u8 *sl = malloc(n);
	u32 *ml = malloc((size_t)n * sizeof(u32));
	signed char *v = malloc(n);
	const u8 **ap = malloc((size_t)n * sizeof(*ap));
	u16 *al = malloc((size_t)n * sizeof(u16));
	if (!keys || !sigs || !msgs || !sp || !mp || !pk || !sl || !ml || !v || !ap || !al) return 1;
	for (u32 i = 0; i < n; i++) {
		keys[i] = kp[i % POOL].pub_key;
		memcpy(sigs + (size_t)i * siglen, psig[i % POOL], siglen);
		memcpy(msgs + (size_t)i * ML, pmsg[i % POOL], ML);
		if (bad_every && i % bad_every == 13) sigs[(size_t)i * siglen + siglen - 3] ^= 2; /* s (BIGN: s1) corrupted */
		sp[i] = sigs + (size_t)i * siglen;
		mp[i] = msgs + (size_t)i * ML;
		pk[i] = &keys[i];
		sl[i] = siglen;
		ml[i] = ML;
		ap[i] = ad;
		al[i] = adl;
	}
	const int has_bad = bad_every && n > 13;
	const unsigned long long msm0 = msm_batches();
	double best = 1e9;
	for (int rep = 0; rep < reps; rep++) {
		double t0 = now_s();
		int r = gpu_vbatch(sp, sl, pk, mp, ml, n, alg, SHA256, ad ? ap : NULL, ad ? al : NULL, NULL, NULL);
		double t = now_s() - t0;
		CHECK(r == (has_bad ? -1 : 0), "batch verdict %d", r);
		if (rep > 0 && t < best) best = t; /* first call builds the comb table and the staging buffers */
		printf("bench rep %d: %.4f s\n", rep, t);
	}
	CHECK(gpu_verdicts(v, n) == n, "verdicts");
	u32 bad = 0;
	for (u32 i = 0; i < n; i++) bad += (v[i] != ((bad_every && i % bad_every == 13) ? -1 : 0));
	CHECK(bad == 0, "%u verdicts wrong", bad);
	printf("DROPIN_BENCH {\"call\": \"%s\", \"curve\": \"%s\", \"items\": %u, \"invalid_every\": %u, "
	       "\"seconds_best\": %.5f, \"verify_per_s\": %.1f, \"struct_bytes_per_key\": %u, \"wrong_verdicts\": %u, "
	       "\"batches_settled_by_multi_scalar_multiplication\": %llu}\n",
	       fn, curve, n, bad_every, best, (double)n / best, (unsigned)sizeof(ec_pub_key), bad, msm_batches() - msm0);
	return failures != 0;
}

/* preload mode: the reference's own high-level code, with prj_pt_mul interposed by the GPU drop-in */
static int run_preload(void)
{
	count_fn calls = (count_fn)dlsym(RTLD_DEFAULT, "eccb200_dropin_call_count");
	if (!calls) {
		printf("FAIL: drop-in not preloaded (eccb200_dropin_call_count missing)\n");
		return 1;
	}
	unsigned long long c0 = calls();
	const char *names[8] = { "SECP256R1", "FRP256V1", "SECP384R1", "BRAINPOOLP256R1", "SECP256K1", "SECP521R1",
				 "SECP224R1", "BRAINPOOLP512R1" };
	const char *only = getenv("HARNESS_CURVES"); /* optional subset, as in direct mode */
	unsigned ncurves = 0;
	for (int c = 0; c < 8; c++) {
		ec_params params;
		if (only && !strstr(only, names[c])) continue;
		ncurves++;
		CHECK(!load_params(&params, names[c]), "import_params");
		u8 qlen = (u8)BYTECEIL(params.ec_gen_order_bitlen);
		hash_alg_type ht = (c == 2) ? SHA384 : ((c == 5 || c == 7) ? SHA512 : SHA256);
		for (int t = 0; t < 6; t++) {
			ec_key_pair kp;
			u8 sig[2 * 66], msg[32];
			CHECK(!ec_key_pair_gen(&kp, &params, ECDSA), "keygen through the GPU");
			for (int j = 0; j < 32; j++) msg[j] = rnd8();
			CHECK(!ec_sign(sig, (u8)(2 * qlen), &kp, msg, 32, ECDSA, ht, NULL, 0), "ec_sign through the GPU");
			CHECK(!ec_verify(sig, (u8)(2 * qlen), &kp.pub_key, msg, 32, ECDSA, ht, NULL, 0), "ec_verify through the GPU");
			sig[1] ^= 1;
			CHECK(ec_verify(sig, (u8)(2 * qlen), &kp.pub_key, msg, 32, ECDSA, ht, NULL, 0) == -1, "forged signature accepted");
		}
		/* ECC-CDH both ways must agree (two variable-base multiplications through the GPU) */
		ec_key_pair a, b;
		u8 sa[66], sb[66], pa[2 * 66], pb[2 * 66], plen = (u8)BYTECEIL(params.ec_fp.p_bitlen);
		CHECK(!ecccdh_gen_key_pair(&a, &params) && !ecccdh_gen_key_pair(&b, &params), "ecccdh keygen");
		CHECK(!ecccdh_serialize_pub_key(&a.pub_key, pa, (u8)(2 * plen)) && !ecccdh_serialize_pub_key(&b.pub_key, pb, (u8)(2 * plen)), "serialize");
		CHECK(!ecccdh_derive_secret(&a.priv_key, pb, (u8)(2 * plen), sa, plen), "derive a");
		CHECK(!ecccdh_derive_secret(&b.priv_key, pa, (u8)(2 * plen), sb, plen), "derive b");
		CHECK(!memcmp(sa, sb, plen), "%s: ECC-CDH secrets differ", names[c]);
	}
	unsigned long long used = calls() - c0;
	count_fn vcount = (count_fn)dlsym(RTLD_DEFAULT, "eccb200_dropin_verify_count");
	unsigned long long verified = vcount ? vcount() : 0;
	printf("preload: %llu prj_pt_mul calls and %llu ec_verify calls served by the GPU drop-in\n", used, verified);
	/* per curve: 6 x (keygen + sign) + 2 + 2 ECC-CDH multiplications through prj_pt_mul; the 12 ec_verify calls are
	 * whole-kernel verifications of the interposed ec_verify */
	CHECK(used >= ncurves * (6 * 2 + 4), "too few interposed calls (%llu): the reference did not go through the drop-in", used);
	CHECK(verified >= ncurves * 12, "ec_verify was not served by the drop-in (%llu)", verified);
	return failures != 0;
}

int main(int argc, char **argv)
{
	int rc;
	if (argc >= 2 && !strcmp(argv[1], "preload")) rc = run_preload();
	else if (argc >= 3 && !strcmp(argv[1], "direct")) rc = run_direct(argv[2]);
	else if (argc >= 3 && !strcmp(argv[1], "threads")) rc = run_threads(argv[2]);
	else if (argc >= 3 && !strcmp(argv[1], "kats")) rc = run_kats(argv[2]);
	else if (argc >= 5 && !strcmp(argv[1], "fuzzmul")) rc = run_fuzzmul(argv[2], argv[3], (u32)strtoul(argv[4], NULL, 10));
	else if (argc >= 5 && !strcmp(argv[1], "fuzz")) rc = run_fuzz(argv[2], argv[3], (u32)strtoul(argv[4], NULL, 10), argc >= 6 ? argv[5] : "SHA256");
	else if (argc >= 5 && !strcmp(argv[1], "bench"))
		rc = run_bench(argv[2], argv[3], (u32)strtoul(argv[4], NULL, 10), argc >= 6 ? argv[5] : "ECDSA",
			       argc >= 7 ? (u32)strtoul(argv[6], NULL, 10) : 64);
	else {
		printf("usage: %s direct <dropin.so> | kats <dropin.so> | fuzz <dropin.so> <curve> <mutants per scheme> [hash name] | fuzzmul <dropin.so> <curve> <iterations> | threads <dropin.so> | bench <dropin.so> <curve> <items> [ECDSA|ECFSDSA|BIP0340|ECSDSA|ECOSDSA|ECKCDSA|ECGDSA|ECRDSA|SM2|BIGN [invalid_every, 0 = none]] | preload\n", argv[0]);
		return 2;
	}
	printf(rc ? "HARNESS FAILED (%d failures)\n" : "HARNESS OK (%d failures)\n", failures);
	return rc;
}
