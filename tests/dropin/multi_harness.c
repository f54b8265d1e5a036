/*
 * tests/dropin/multi_harness.c — a plain-C host of the multi-device C ABI (include/libecc_b200.h: eccb200_multi_*).
 *
 * One process, one call: the library shards the batch over the devices named on the command line.  The results are
 * judged by the UNMODIFIED reference linked into this binary (oracle/_ref/libecc_ref.so): prj_pt_mul on the curve's
 * generator + prj_pt_export_to_aff_buf for sampled items of every shard, ec_sign / ec_verify for the signatures.
 * Built by oracle/Makefile in the container that has the reference's headers; the binary travels with oracle/_ref/.
 *
 *   multi_harness <path/to/libecc_b200.so> <device list, e.g. 0,1,2,3 or 0,0> [items]
 */
#define _GNU_SOURCE
#include "libsig.h"
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct eccb200_multi eccb200_multi;
typedef int (*create_fn)(eccb200_multi **, int, const int *, int, int);
typedef void (*destroy_fn)(eccb200_multi *);
typedef int (*count_fn)(const eccb200_multi *);
typedef int (*mul_fn)(eccb200_multi *, uint64_t, const uint8_t *, const uint8_t *, uint8_t *, int8_t *);
typedef int (*verify_fn)(eccb200_multi *, uint64_t, const uint8_t *, const uint8_t *, const uint8_t *, uint32_t, int8_t *);
typedef void *(*halloc_fn)(size_t);
typedef void (*hfree_fn)(void *);
typedef const char *(*err_fn)(void);

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

static unsigned long long rng_state = 0x6c69626563632d32ULL;
static u8 rnd8(void)
{
	rng_state += 0x9e3779b97f4a7c15ULL;
	unsigned long long z = rng_state;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return (u8)((z ^ (z >> 31)) >> 24);
}

int main(int argc, char **argv)
{
	if (argc < 3) {
		printf("usage: %s libecc_b200.so dev[,dev...] [items]\n", argv[0]);
		return 2;
	}
	void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
	if (!h) {
		printf("dlopen: %s\n", dlerror());
		return 2;
	}
	create_fn m_create = (create_fn)dlsym(h, "eccb200_multi_create");
	destroy_fn m_destroy = (destroy_fn)dlsym(h, "eccb200_multi_destroy");
	count_fn m_count = (count_fn)dlsym(h, "eccb200_multi_device_count");
	mul_fn m_mul = (mul_fn)dlsym(h, "eccb200_multi_prj_pt_mul_batch");
	verify_fn m_verify = (verify_fn)dlsym(h, "eccb200_multi_ecdsa_verify_batch");
	halloc_fn h_alloc = (halloc_fn)dlsym(h, "eccb200_host_alloc");
	hfree_fn h_free = (hfree_fn)dlsym(h, "eccb200_host_free");
	err_fn last_err = (err_fn)dlsym(h, "eccb200_last_error");
	if (!m_create || !m_destroy || !m_count || !m_mul || !m_verify || !h_alloc || !h_free || !last_err) {
		printf("missing symbol\n");
		return 2;
	}
	int devs[64], ndev = 0;
	for (char *tok = strtok(argv[2], ","); tok && ndev < 64; tok = strtok(NULL, ",")) devs[ndev++] = atoi(tok);
	const uint64_t n = argc > 3 ? strtoull(argv[3], NULL, 10) : 100003; /* odd on purpose: ragged shards */

	const struct { const char *name; int id; } curves[] = { { "SECP256R1", 4 }, { "SECP384R1", 5 } };
	for (unsigned c = 0; c < sizeof(curves) / sizeof(curves[0]); c++) {
		ec_params params;
		const ec_str_params *sp = NULL;
		CHECK(!ec_get_curve_params_by_name((const u8 *)curves[c].name, (u8)(strlen(curves[c].name) + 1), &sp) && sp,
		      "curve lookup");
		CHECK(!import_params(&params, sp), "import_params");
		const u32 plen = (u32)BYTECEIL(params.ec_fp.p_bitlen), qlen = (u32)BYTECEIL(params.ec_gen_order_bitlen);
		eccb200_multi *m = NULL;
		if (m_create(&m, curves[c].id, devs, ndev, 16)) {
			printf("eccb200_multi_create: %s\n", last_err());
			return 1;
		}
		CHECK(m_count(m) == ndev, "device count");

		/* ---- fixed base, page-locked and pageable buffers, ragged shards */
		for (int pinned = 1; pinned >= 0; pinned--) {
			u8 *sc = pinned ? h_alloc(n * qlen) : malloc(n * qlen);
			u8 *out = pinned ? h_alloc(n * 2 * plen) : malloc(n * 2 * plen);
			int8_t *st = pinned ? h_alloc(n) : malloc(n);
			for (uint64_t i = 0; i < n * qlen; i++) sc[i] = rnd8();
			memset(sc, 0, qlen);                       /* k = 0 -> infinity */
			memset(sc + (n - 1) * qlen, 0xff, qlen);   /* k = 2^l - 1: reduced mod q */
			memset(out, 0xaa, n * 2 * plen);
			memset(st, 0x55, n);
			if (m_mul(m, n, sc, NULL, out, st)) {
				printf("eccb200_multi_prj_pt_mul_batch: %s\n", last_err());
				return 1;
			}
			/* the reference on a sample: the first and last items of every shard plus a stride through the batch */
			uint64_t checked = 0;
			for (uint64_t i = 0; i < n; i++) {
				int edge = 0;
				for (int g = 0; g <= ndev; g++) {
					uint64_t b = n * (uint64_t)g / (uint64_t)ndev;
					if (i == b || i + 1 == b || i == b + 1) edge = 1;
				}
				if (!edge && i % 997) continue;
				nn k;
				prj_pt r;
				u8 want[2 * 72];
				CHECK(!nn_init_from_buf(&k, sc + i * qlen, (u16)qlen), "nn_init_from_buf");
				CHECK(!prj_pt_mul(&r, &k, &params.ec_gen), "reference prj_pt_mul");
				int iszero = 0;
				CHECK(!prj_pt_iszero(&r, &iszero), "prj_pt_iszero");
				if (iszero) {
					CHECK(st[i] == 1, "%s item %llu: status %d, reference says infinity", curves[c].name,
					      (unsigned long long)i, st[i]);
				} else {
					CHECK(!prj_pt_export_to_aff_buf(&r, want, 2 * plen), "export");
					CHECK(st[i] == 0 && !memcmp(want, out + i * 2 * plen, 2 * plen),
					      "%s item %llu (pinned=%d): differs from the reference", curves[c].name,
					      (unsigned long long)i, pinned);
				}
				checked++;
			}
			uint64_t untouched = 0;
			for (uint64_t i = 0; i < n; i++) untouched += (st[i] == 0x55);
			CHECK(untouched == 0, "%llu status bytes never written", (unsigned long long)untouched);
			printf("%s fixed base, %d device contexts, %llu items (%s buffers): %llu checked against the reference\n",
			       curves[c].name, ndev, (unsigned long long)n, pinned ? "page-locked" : "pageable",
			       (unsigned long long)checked);
			if (pinned) { h_free(sc); h_free(out); h_free(st); } else { free(sc); free(out); free(st); }
		}

		/* ---- ECDSA verification on digests: signatures made and judged by the reference */
		{
			enum { NS = 301 };
			const hash_alg_type ht = (plen == 32) ? SHA256 : SHA384;
			const u32 hlen = (plen == 32) ? 32 : 48;
			u8 *sigs = malloc((size_t)NS * 2 * qlen), *pubs = malloc((size_t)NS * 2 * plen), *dg = malloc((size_t)NS * hlen);
			int8_t *v = malloc(NS), *want = malloc(NS);
			for (int i = 0; i < NS; i++) {
				ec_key_pair kp;
				u8 msg[40];
				for (int j = 0; j < 40; j++) msg[j] = rnd8();
				CHECK(!ec_key_pair_gen(&kp, &params, ECDSA), "ec_key_pair_gen");
				CHECK(!ec_sign(sigs + (size_t)i * 2 * qlen, (u8)(2 * qlen), &kp, msg, 40, ECDSA, ht, NULL, 0), "ec_sign");
				if (i % 7 == 3) sigs[(size_t)i * 2 * qlen + 5] ^= 0x40; /* corrupt r */
				if (i % 11 == 5) msg[0] ^= 1;                           /* other message */
				CHECK(!ec_pub_key_export_to_aff_buf(&kp.pub_key, pubs + (size_t)i * 2 * plen, (u8)(2 * plen)), "pub export");
				const hash_mapping *hm = NULL;
				CHECK(!get_hash_by_type(ht, &hm) && hm, "hash mapping");
				const u8 *in[2] = { msg, NULL };
				u32 il[1] = { 40 };
				CHECK(!hm->hfunc_scattered(in, il, dg + (size_t)i * hlen), "hash");
				want[i] = ec_verify(sigs + (size_t)i * 2 * qlen, (u8)(2 * qlen), &kp.pub_key, msg, 40, ECDSA, ht, NULL, 0) ? -1 : 0;
			}
			if (m_verify(m, NS, sigs, pubs, dg, hlen, v)) {
				printf("eccb200_multi_ecdsa_verify_batch: %s\n", last_err());
				return 1;
			}
			int bad = 0, rejected = 0;
			for (int i = 0; i < NS; i++) {
				bad += v[i] != want[i];
				rejected += want[i] != 0;
			}
			CHECK(bad == 0, "%s: %d verdicts differ from the reference's ec_verify", curves[c].name, bad);
			CHECK(rejected > 0 && rejected < NS, "corruption pattern did not produce both verdicts");
			printf("%s ECDSA verify over %d device contexts: %d signatures, %d rejected, all verdicts equal the reference's\n",
			       curves[c].name, ndev, NS, rejected);
			free(sigs); free(pubs); free(dg); free(v); free(want);
		}
		m_destroy(m);
	}
	if (failures) {
		printf("HARNESS FAILED: %d\n", failures);
		return 1;
	}
	printf("HARNESS OK\n");
	return 0;
}
