/*
 * oracle/ecc_oracle.c — plain-C CPU restatement of the reference's (ANSSI-FR/libecc) prj_pt_mul and ECDSA
 * verification path.  TEST INFRASTRUCTURE ONLY (see ecc_oracle.h).  Parity status: PINNED (see ecc_oracle.h).
 *
 * Every function cites the reference code it follows (paths relative to /root/reference/src).  The structure
 * is deliberately the reference's: 64-bit little-endian limbs, CIOS Montgomery product, Renes-Costello-Batina
 * complete addition, MSB-fixed Montgomery ladder that doubles with the addition law, Fermat inversion.
 *
 * Deliberate, result-preserving simplifications (all invisible at the affine level, which is the parity level
 * SURVEY.md §8c defines):
 *   - no projective blinding (prj_pt.c:1266-1291) and no address masking r (prj_pt.c:1631): lambda = 1, r = 0;
 *   - bare u64[n] instead of the 27-word nn container with magic/wlen (nn.h:67-71);
 *   - WORD_MUL (words.h:98-127) is one unsigned __int128 product;
 *   - s^-1 mod q by Fermat instead of the binary xgcd of nn_modinv.c:220 (same unique inverse, q prime);
 *   - Montgomery constants (mpinv, R, R^2) are derived here instead of read from the curves/known headers.
 */
#include "ecc_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

#define MAXL 9   /* limbs of a field element / order (4 for 256-bit, 6 for 384-bit, 9 for 521-bit) */
#define BIGL 24  /* capacity for scalars (reference: nn up to 27 words, nn_config.h:154) */

static __thread u64 g_mul_count;
static __thread u64 g_last_mul_count;

/* ------------------------------------------------------------------------------------------------ nn level */

/* restates nn_cmp (nn/nn.c:360) on fixed n limbs: -1, 0, 1 */
static int nn_cmp_n(const u64 *a, const u64 *b, int n)
{
	for (int i = n - 1; i >= 0; i--) {
		if (a[i] != b[i]) return (a[i] < b[i]) ? -1 : 1;
	}
	return 0;
}

static int nn_iszero_n(const u64 *a, int n)
{
	u64 acc = 0;
	for (int i = 0; i < n; i++) acc |= a[i];
	return acc == 0;
}

/* out = a + b, returns carry (nn_add, nn/nn_add.c:167) */
static u64 nn_add_n(u64 *out, const u64 *a, const u64 *b, int n)
{
	u64 carry = 0;
	for (int i = 0; i < n; i++) {
		u128 t = (u128)a[i] + b[i] + carry;
		out[i] = (u64)t;
		carry = (u64)(t >> 64);
	}
	return carry;
}

/* out = a - b, returns borrow (nn_sub, nn/nn_add.c:291) */
static u64 nn_sub_n(u64 *out, const u64 *a, const u64 *b, int n)
{
	u64 borrow = 0;
	for (int i = 0; i < n; i++) {
		u128 t = (u128)a[i] - b[i] - borrow;
		out[i] = (u64)t;
		borrow = (u64)(t >> 64) & 1;
	}
	return borrow;
}

/* nn_bitlen (nn/nn_logical.c:514) */
static int nn_bitlen_n(const u64 *a, int n)
{
	for (int i = n - 1; i >= 0; i--) {
		if (a[i]) return 64 * i + (64 - __builtin_clzll(a[i]));
	}
	return 0;
}

/* nn_getbit (nn/nn_logical.c:541) */
static int nn_getbit_n(const u64 *a, int bit)
{
	return (int)((a[bit / 64] >> (bit % 64)) & 1);
}

/* nn_init_from_buf (nn/nn.c:479): big-endian bytes -> little-endian limbs */
static void nn_from_be(u64 *out, int n, const uint8_t *buf, uint32_t len)
{
	memset(out, 0, sizeof(u64) * (size_t)n);
	for (uint32_t i = 0; i < len; i++) {
		uint32_t pos = len - 1 - i; /* byte significance */
		if (pos / 8 < (uint32_t)n) out[pos / 8] |= (u64)buf[i] << (8 * (pos % 8));
	}
}

/* nn_export_to_buf (nn/nn.c:511) */
static void nn_to_be(uint8_t *buf, uint32_t len, const u64 *in, int n)
{
	for (uint32_t i = 0; i < len; i++) {
		uint32_t pos = len - 1 - i;
		buf[i] = (pos / 8 < (uint32_t)n) ? (uint8_t)(in[pos / 8] >> (8 * (pos % 8))) : 0;
	}
}

/* Montgomery context: restates fp_ctx (fp/fp.h:31-57) — p, mpinv = -p^-1 mod 2^64, r = R mod p, r2 = R^2 mod p */
typedef struct {
	int n;
	int bitlen;
	u64 p[MAXL];
	u64 mpinv;
	u64 r[MAXL];
	u64 r2[MAXL];
} mctx;

/*
 * CIOS Montgomery product, out = a*b*R^-1 mod p.  Restates _nn_mul_redc1 (nn/nn_mul_redc1.c:124-218):
 * per outer iteration i: n multiply-accumulates with carry (:175-183), carry tail into word n and the extra
 * bit 'acc' (:184-189), m = out[0]*mpinv (:191), n multiply-accumulates with p merged with the one-word
 * right shift (:192-204); final conditional subtraction (:210-211).  out must not alias a or b.
 */
static void mul_redc1(u64 *out, const u64 *a, const u64 *b, const mctx *c)
{
	int n = c->n;
	u64 t[MAXL + 1];
	u64 carry, acc, m;
	g_mul_count++;
	for (int i = 0; i <= n; i++) t[i] = 0;
	for (int i = 0; i < n; i++) {
		carry = 0;
		for (int j = 0; j < n; j++) {
			u128 pr = (u128)a[i] * b[j] + t[j] + carry;
			t[j] = (u64)pr;
			carry = (u64)(pr >> 64);
		}
		{
			u128 s = (u128)t[n] + carry;
			t[n] = (u64)s;
			acc = (u64)(s >> 64);
		}
		m = t[0] * c->mpinv;
		{
			u128 pr = (u128)m * c->p[0] + t[0];
			carry = (u64)(pr >> 64);
		}
		for (int j = 1; j < n; j++) {
			u128 pr = (u128)m * c->p[j] + t[j] + carry;
			t[j - 1] = (u64)pr;
			carry = (u64)(pr >> 64);
		}
		{
			u128 s = (u128)t[n] + carry;
			t[n - 1] = (u64)s;
			t[n] = acc + (u64)(s >> 64);
		}
	}
	/* msw is 0 or 1; subtract p if t >= p */
	if (t[n] || nn_cmp_n(t, c->p, n) >= 0) nn_sub_n(t, t, c->p, n);
	memcpy(out, t, sizeof(u64) * (size_t)n);
}

/* nn_mod_add (nn/nn_add.c:337): inputs < p */
static void mod_add(u64 *out, const u64 *a, const u64 *b, const mctx *c)
{
	u64 t[MAXL];
	u64 carry = nn_add_n(t, a, b, c->n);
	if (carry || nn_cmp_n(t, c->p, c->n) >= 0) nn_sub_n(t, t, c->p, c->n);
	memcpy(out, t, sizeof(u64) * (size_t)c->n);
}

/* nn_mod_sub (nn/nn_add.c:398): inputs < p; add p back when a < b */
static void mod_sub(u64 *out, const u64 *a, const u64 *b, const mctx *c)
{
	u64 t[MAXL];
	u64 borrow = nn_sub_n(t, a, b, c->n);
	if (borrow) nn_add_n(t, t, c->p, c->n);
	memcpy(out, t, sizeof(u64) * (size_t)c->n);
}

/* nn_compute_redc1_coefs (nn/nn_mul_redc1.c:40): mpinv by Newton iteration, R and R^2 by repeated doubling */
static void mctx_init(mctx *c, const u64 *p, int n)
{
	u64 inv = 1;
	c->n = n;
	memcpy(c->p, p, sizeof(u64) * (size_t)n);
	c->bitlen = nn_bitlen_n(p, n);
	for (int i = 0; i < 6; i++) inv *= 2 - p[0] * inv; /* p^-1 mod 2^64 */
	c->mpinv = (u64)0 - inv;
	/* r = 2^(64n) mod p: start from 1, double 64n times */
	u64 x[MAXL];
	memset(x, 0, sizeof(x));
	x[0] = 1;
	for (int i = 0; i < 64 * n; i++) mod_add(x, x, x, c);
	memcpy(c->r, x, sizeof(x));
	for (int i = 0; i < 64 * n; i++) mod_add(x, x, x, c);
	memcpy(c->r2, x, sizeof(x));
}

/* fp_redcify / fp_unredcify (fp/fp_mul_redc1.c:62,79) */
static void to_monty(u64 *out, const u64 *a, const mctx *c)
{
	u64 t[MAXL];
	mul_redc1(t, a, c->r2, c);
	memcpy(out, t, sizeof(u64) * (size_t)c->n);
}

static void from_monty(u64 *out, const u64 *a, const mctx *c)
{
	u64 one[MAXL], t[MAXL];
	memset(one, 0, sizeof(one));
	one[0] = 1;
	mul_redc1(t, a, one, c);
	memcpy(out, t, sizeof(u64) * (size_t)c->n);
}

/* fp_mul (fp/fp_mul.c:23-40): plain modular product a*b mod p.  The reference does nn_mul + nn_mod_unshifted;
 * the canonical result is reproduced here as redc(redc(a,b), R^2). */
static void mod_mul(u64 *out, const u64 *a, const u64 *b, const mctx *c)
{
	u64 t[MAXL], u[MAXL];
	mul_redc1(t, a, b, c);
	mul_redc1(u, t, c->r2, c);
	memcpy(out, u, sizeof(u64) * (size_t)c->n);
}

/* fp_inv (fp/fp_mul.c:51-68) -> nn_modinv_fermat_redc (nn/nn_modinv.c:538) -> nn_mod_pow_redc
 * (nn/nn_mod_pow.c:39-155): out = a^(p-2) mod p by a left-to-right ladder in the Montgomery domain. */
static void mod_inv_fermat(u64 *out, const u64 *a, const mctx *c)
{
	u64 e[MAXL], two[MAXL], am[MAXL], acc[MAXL], t[MAXL];
	memset(two, 0, sizeof(two));
	two[0] = 2;
	nn_sub_n(e, c->p, two, c->n);
	to_monty(am, a, c);
	memcpy(acc, c->r, sizeof(acc)); /* 1 in Montgomery form */
	for (int i = nn_bitlen_n(e, c->n) - 1; i >= 0; i--) {
		mul_redc1(t, acc, acc, c);
		memcpy(acc, t, sizeof(t));
		if (nn_getbit_n(e, i)) {
			mul_redc1(t, acc, am, c);
			memcpy(acc, t, sizeof(t));
		}
	}
	from_monty(out, acc, c);
}

/* ------------------------------------------------------------------------------------------------ curves */

typedef struct {
	const char *name;
	int n;          /* limbs */
	uint32_t plen;  /* bytes of p */
	uint32_t qlen;  /* bytes of q */
	int qbits;
	mctx fp;        /* mod p */
	mctx fq;        /* mod q */
	u64 a[MAXL], b[MAXL];              /* ec_shortw_crv.a / .b  (curves/ec_shortw.h:25-36) */
	u64 a_monty[MAXL], b3_monty[MAXL]; /* a*R, 3b*R (curves/ec_shortw.c:75,84-87) */
	u64 gx[MAXL], gy[MAXL];
} curve_t;

typedef struct {
	u64 X[MAXL], Y[MAXL], Z[MAXL]; /* prj_pt (curves/prj_pt.h:26-32), coordinates in normal form */
} pt_t;

static const struct {
	const char *name;
	const char *p, *q, *a, *b, *gx, *gy;
} CURVE_HEX[] = {
	/* values as published (FIPS 186-4 D.1.2.3 / D.1.2.4, ANSSI JORF 2011 FRP256v1); tests/test_oracle.py checks
	 * them against the reference's curves/known/ec_params_*.h through ref_curve_info(). */
	{ "SECP256R1",
	  "ffffffff00000001000000000000000000000000ffffffffffffffffffffffff",
	  "ffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551",
	  "ffffffff00000001000000000000000000000000fffffffffffffffffffffffc",
	  "5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b",
	  "6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296",
	  "4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5" },
	{ "FRP256V1",
	  "f1fd178c0b3ad58f10126de8ce42435b3961adbcabc8ca6de8fcf353d86e9c03",
	  "f1fd178c0b3ad58f10126de8ce42435b53dc67e140d2bf941ffdd459c6d655e1",
	  "f1fd178c0b3ad58f10126de8ce42435b3961adbcabc8ca6de8fcf353d86e9c00",
	  "ee353fca5428a9300d4aba754a44c00fdfec0c9ae4b1a1803075ed967b7bb73f",
	  "b6b3d4c356c139eb31183d4749d423958c27d2dcaf98b70164c97a2dd98f5cff",
	  "6142e0f7c8b204911f9271f0f3ecef8c2701c307e8e4c9e183115a1554062cfb" },
	{ "SECP384R1",
	  "fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffeffffffff0000000000000000ffffffff",
	  "ffffffffffffffffffffffffffffffffffffffffffffffffc7634d81f4372ddf581a0db248b0a77aecec196accc52973",
	  "fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffeffffffff0000000000000000fffffffc",
	  "b3312fa7e23ee7e4988e056be3f82d19181d9c6efe8141120314088f5013875ac656398d8a2ed19d2a85c8edd3ec2aef",
	  "aa87ca22be8b05378eb1c71ef320ad746e1d3b628ba79b9859f741e082542a385502f25dbf55296c3a545e3872760ab7",
	  "3617de4a96262c6f5d9e98bf9292dc29f8f41dbd289a147ce9da3113b5f0b8c00a60b1ce1d7e819d7a431d7c90ea0e5f" },
	/* additional curves (RFC 5639 Brainpool, SEC 2 secp256k1): same code path, general a */
	{ "BRAINPOOLP256R1",
	  "a9fb57dba1eea9bc3e660a909d838d726e3bf623d52620282013481d1f6e5377",
	  "a9fb57dba1eea9bc3e660a909d838d718c397aa3b561a6f7901e0e82974856a7",
	  "7d5a0975fc2c3057eef67530417affe7fb8055c126dc5c6ce94a4b44f330b5d9",
	  "26dc5c6ce94a4b44f330b5d9bbd77cbf958416295cf7e1ce6bccdc18ff8c07b6",
	  "8bd2aeb9cb7e57cb2c4b482ffc81b7afb9de27e1e3bd23c23a4453bd9ace3262",
	  "547ef835c3dac4fd97f8461a14611dc9c27745132ded8e545c1d54c72f046997" },
	{ "BRAINPOOLP384R1",
	  "8cb91e82a3386d280f5d6f7e50e641df152f7109ed5456b412b1da197fb71123acd3a729901d1a71874700133107ec53",
	  "8cb91e82a3386d280f5d6f7e50e641df152f7109ed5456b31f166e6cac0425a7cf3ab6af6b7fc3103b883202e9046565",
	  "7bc382c63d8c150c3c72080ace05afa0c2bea28e4fb22787139165efba91f90f8aa5814a503ad4eb04a8c7dd22ce2826",
	  "04a8c7dd22ce28268b39b55416f0447c2fb77de107dcd2a62e880ea53eeb62d57cb4390295dbc9943ab78696fa504c11",
	  "1d1c64f068cf45ffa2a63a81b7c13f6b8847a3e77ef14fe3db7fcafe0cbd10e8e826e03436d646aaef87b2e247d4af1e",
	  "8abe1d7520f9c2a45cb1eb8e95cfd55262b70b29feec5864e19c054ff99129280e4646217791811142820341263c5315" },
	{ "SECP256K1",
	  "fffffffffffffffffffffffffffffffffffffffffffffffffffffffefffffc2f",
	  "fffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364141",
	  "0000000000000000000000000000000000000000000000000000000000000000",
	  "0000000000000000000000000000000000000000000000000000000000000007",
	  "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798",
	  "483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8" },
	/* 521-bit: 66-byte strings, 9 limbs (curves/known/ec_params_secp521r1.h) */
	{ "SECP521R1",
	  "01ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff",
	  "01fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffa51868783bf2f966b7fcc0148f709a5d03bb5c9b8899c47aebb6fb71e91386409",
	  "01fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffc",
	  "0051953eb9618e1c9a1f929a21a0b68540eea2da725b99b315f3b8b489918ef109e156193951ec7e937b1652c0bd3bb1bf073573df883d2c34f1ef451fd46b503f00",
	  "00c6858e06b70404e9cd9e3ecb662395b4429c648139053fb521f828af606b4d3dbaa14b5e77efe75928fe1dc127a2ffa8de3348b3c1856a429bf97e7e31c2e5bd66",
	  "011839296a789a3bc0045c8a5fb42c7d1bd998f54449579b446817afbd17273e662c97ee72995ef42640c550b9013fad0761353c7086a272c24088be94769fd16650" },
	{ "SM2P256V1",
	  "fffffffeffffffffffffffffffffffffffffffff00000000ffffffffffffffff",
	  "fffffffeffffffffffffffffffffffff7203df6b21c6052b53bbf40939d54123",
	  "fffffffeffffffffffffffffffffffffffffffff00000000fffffffffffffffc",
	  "28e9fa9e9d9f5e344d5a9e4bcf6509a7f39789f515ab8f92ddbcbd414d940e93",
	  "32c4ae2c1f1981195f9904466a39c9948fe30bbff2660be1715a4589334c74c7",
	  "bc3736a2f4f6779c59bdcee36b692153d0a9877cc62a474002df32e52139f0a0" },
	{ "BRAINPOOLP512R1",
	  "aadd9db8dbe9c48b3fd4e6ae33c9fc07cb308db3b3c9d20ed6639cca703308717d4d9b009bc66842aecda12ae6a380e62881ff2f2d82c68528aa6056583a48f3",
	  "aadd9db8dbe9c48b3fd4e6ae33c9fc07cb308db3b3c9d20ed6639cca70330870553e5c414ca92619418661197fac10471db1d381085ddaddb58796829ca90069",
	  "7830a3318b603b89e2327145ac234cc594cbdd8d3df91610a83441caea9863bc2ded5d5aa8253aa10a2ef1c98b9ac8b57f1117a72bf2c7b9e7c1ac4d77fc94ca",
	  "3df91610a83441caea9863bc2ded5d5aa8253aa10a2ef1c98b9ac8b57f1117a72bf2c7b9e7c1ac4d77fc94cadc083e67984050b75ebae5dd2809bd638016f723",
	  "81aee4bdd82ed9645a21322e9c4c6a9385ed9f70b5d916c1b43b62eef4d0098eff3b1f78e2d0d48d50d1687b93b97d5f7c6d5047406a5e688b352209bcb9f822",
	  "7dde385d566332ecc0eabfa9cf7822fdf209f70024a57b1aa000c55b881f8111b2dcde494a5f485e5bca4bd88a2763aed1ca2b2fa8f0540678cd1e0f3ad80892" },
	{ "SECP224R1",
	  "ffffffffffffffffffffffffffffffff000000000000000000000001",
	  "ffffffffffffffffffffffffffff16a2e0b8f03e13dd29455c5c2a3d",
	  "fffffffffffffffffffffffffffffffefffffffffffffffffffffffe",
	  "b4050a850c04b3abf54132565044b0b7d7bfd8ba270b39432355ffb4",
	  "b70e0cbd6bb4bf7f321390b94a03c1d356c21122343280d6115c1d21",
	  "bd376388b5f723fb4c22dfe6cd4375a05a07476444d5819985007e34" },
	{ "SECP192R1",
	  "fffffffffffffffffffffffffffffffeffffffffffffffff",
	  "ffffffffffffffffffffffff99def836146bc9b1b4d22831",
	  "fffffffffffffffffffffffffffffffefffffffffffffffc",
	  "64210519e59c80e70fa7e9ab72243049feb8deecc146b9b1",
	  "188da80eb03090f67cbf20eb43a18800f4ff0afd82ff1012",
	  "07192b95ffc8da78631011ed6b24cdd573f977a11e794811" },
};

static void hex_to_limbs(u64 *out, int n, const char *hex)
{
	size_t len = strlen(hex);
	memset(out, 0, sizeof(u64) * (size_t)n);
	for (size_t i = 0; i < len; i++) {
		char ch = hex[len - 1 - i];
		u64 v = (u64)((ch >= 'a') ? (ch - 'a' + 10) : (ch - '0'));
		out[i / 16] |= v << (4 * (i % 16));
	}
}

/* import_params (curves/ec_params.c:24-194) + ec_shortw_crv_init (curves/ec_shortw.c:41-97) */
static int curve_load(curve_t *c, const char *name)
{
	for (size_t k = 0; k < sizeof(CURVE_HEX) / sizeof(CURVE_HEX[0]); k++) {
		if (strcmp(CURVE_HEX[k].name, name)) continue;
		u64 p[MAXL], q[MAXL], b3[MAXL];
		size_t hl = strlen(CURVE_HEX[k].p);
		c->name = CURVE_HEX[k].name;
		c->n = (int)((hl + 15) / 16);
		c->plen = (uint32_t)(hl / 2);
		c->qlen = (uint32_t)(strlen(CURVE_HEX[k].q) / 2);
		hex_to_limbs(p, c->n, CURVE_HEX[k].p);
		hex_to_limbs(q, c->n, CURVE_HEX[k].q);
		mctx_init(&c->fp, p, c->n);
		mctx_init(&c->fq, q, c->n);
		c->qbits = c->fq.bitlen;
		hex_to_limbs(c->a, c->n, CURVE_HEX[k].a);
		hex_to_limbs(c->b, c->n, CURVE_HEX[k].b);
		hex_to_limbs(c->gx, c->n, CURVE_HEX[k].gx);
		hex_to_limbs(c->gy, c->n, CURVE_HEX[k].gy);
		to_monty(c->a_monty, c->a, &c->fp);
		mod_add(b3, c->b, c->b, &c->fp);
		mod_add(b3, b3, c->b, &c->fp);
		to_monty(c->b3_monty, b3, &c->fp);
		g_mul_count = 0;
		return 0;
	}
	return -1;
}

int ora_curve_sizes(const char *curve, uint32_t *plen, uint32_t *qlen)
{
	curve_t c;
	if (curve_load(&c, curve)) return -1;
	*plen = c.plen;
	*qlen = c.qlen;
	return 0;
}

static int pt_iszero(const pt_t *p, const curve_t *c) { return nn_iszero_n(p->Z, c->n); }

/* prj_pt_is_on_curve (curves/prj_pt.c:144-190): Y^2 Z == X^3 + a X Z^2 + b Z^3, same operation order */
static int pt_is_on_curve(const pt_t *in, const curve_t *c)
{
	const mctx *f = &c->fp;
	u64 X[MAXL], Y[MAXL], Z[MAXL];
	mod_mul(X, in->X, in->X, f);
	mod_mul(X, X, in->X, f);
	mod_mul(Z, in->X, c->a, f);
	mod_mul(Y, c->b, in->Z, f);
	mod_add(Z, Z, Y, f);
	mod_mul(Z, Z, in->Z, f);
	mod_mul(Z, Z, in->Z, f);
	mod_add(X, X, Z, f);
	mod_mul(Y, in->Y, in->Y, f);
	mod_mul(Y, Y, in->Z, f);
	return nn_cmp_n(X, Y, c->n) == 0;
}

/*
 * __prj_pt_add_monty_cf (curves/prj_pt.c:971-1071): Renes-Costello-Batina complete addition (Alg. 1, general a),
 * 17 fp_mul_monty + 17 fp_add + 6 fp_sub in exactly the reference's order, on normal-form coordinates with
 * Montgomery-form constants (homogeneity makes that consistent).  Returns -1 on the Y=Z=0
 * exceptional output (:1058-1060).
 */
static int pt_add_cf(pt_t *out, const pt_t *in1, const pt_t *in2, const curve_t *c)
{
	const mctx *f = &c->fp;
	u64 t0[MAXL], t1[MAXL], t2[MAXL], t3[MAXL], t4[MAXL], t5[MAXL], X3[MAXL], Y3[MAXL], Z3[MAXL], u[MAXL];
	mul_redc1(t0, in1->X, in2->X, f);
	mul_redc1(t1, in1->Y, in2->Y, f);
	mul_redc1(t2, in1->Z, in2->Z, f);
	mod_add(t3, in1->X, in1->Y, f);
	mod_add(t4, in2->X, in2->Y, f);
	mul_redc1(u, t3, t4, f); memcpy(t3, u, sizeof(u));
	mod_add(t4, t0, t1, f);
	mod_sub(t3, t3, t4, f);
	mod_add(t4, in1->X, in1->Z, f);
	mod_add(t5, in2->X, in2->Z, f);
	mul_redc1(u, t4, t5, f); memcpy(t4, u, sizeof(u));
	mod_add(t5, t0, t2, f);
	mod_sub(t4, t4, t5, f);
	mod_add(t5, in1->Y, in1->Z, f);
	mod_add(X3, in2->Y, in2->Z, f);
	mul_redc1(u, t5, X3, f); memcpy(t5, u, sizeof(u));
	mod_add(X3, t1, t2, f);
	mod_sub(t5, t5, X3, f);
	mul_redc1(Z3, c->a_monty, t4, f);
	mul_redc1(X3, c->b3_monty, t2, f);
	mod_add(Z3, X3, Z3, f);
	mod_sub(X3, t1, Z3, f);
	mod_add(Z3, t1, Z3, f);
	mul_redc1(Y3, X3, Z3, f);
	mod_add(t1, t0, t0, f);
	mod_add(t1, t1, t0, f);
	mul_redc1(u, c->a_monty, t2, f); memcpy(t2, u, sizeof(u));
	mul_redc1(u, c->b3_monty, t4, f); memcpy(t4, u, sizeof(u));
	mod_add(t1, t1, t2, f);
	mod_sub(t2, t0, t2, f);
	mul_redc1(u, c->a_monty, t2, f); memcpy(t2, u, sizeof(u));
	mod_add(t4, t4, t2, f);
	mul_redc1(t0, t1, t4, f);
	mod_add(Y3, Y3, t0, f);
	mul_redc1(t0, t5, t4, f);
	mul_redc1(u, t3, X3, f); memcpy(X3, u, sizeof(u));
	mod_sub(X3, X3, t0, f);
	mul_redc1(t0, t3, t1, f);
	mul_redc1(u, t5, Z3, f); memcpy(Z3, u, sizeof(u));
	mod_add(Z3, Z3, t0, f);
	memcpy(out->X, X3, sizeof(X3));
	memcpy(out->Y, Y3, sizeof(Y3));
	memcpy(out->Z, Z3, sizeof(Z3));
	if (nn_iszero_n(Z3, c->n) && nn_iszero_n(Y3, c->n)) return -1;
	return 0;
}

/*
 * _prj_pt_mul_ltr_monty_ladder (curves/prj_pt.c:1569-1720) with r = 0 and lambda = 1.
 *   m' recoding (:1591-1619): m < q -> m+q, plus q again if bitlen(m+q) == bitlen(q);
 *                             q <= m < q^2 -> same with q^2;  m >= q^2 -> m unchanged.
 *   T[0] = in, T[1] = in+in (complete add as doubling, :1654); per bit (:1660-1702):
 *   T[2] = T[mbit]+T[mbit]; T[1] = T[0]+T[1]; T[0] = T[2-mbit]; T[1] = T[1+mbit].
 * m: mlimbs little-endian limbs.
 */
static int pt_mul_ladder(pt_t *out, const u64 *m, int mlimbs, const pt_t *in, const curve_t *c)
{
	u64 mm[BIGL + 2], qq[BIGL + 2], q2[BIGL + 2];
	int n = c->n, W = BIGL + 2;
	int ret_ops = 0;
	memset(mm, 0, sizeof(mm));
	memset(qq, 0, sizeof(qq));
	memset(q2, 0, sizeof(q2));
	memcpy(mm, m, sizeof(u64) * (size_t)mlimbs);
	memcpy(qq, c->fq.p, sizeof(u64) * (size_t)n);
	/* q^2 (nn_sqr, :1588) — schoolbook */
	for (int i = 0; i < n; i++) {
		u64 carry = 0;
		for (int j = 0; j < n; j++) {
			u128 pr = (u128)qq[i] * qq[j] + q2[i + j] + carry;
			q2[i + j] = (u64)pr;
			carry = (u64)(pr >> 64);
		}
		q2[i + n] += carry;
	}
	if (nn_cmp_n(mm, qq, W) < 0) {
		nn_add_n(mm, mm, qq, W);
		if (nn_bitlen_n(mm, W) == nn_bitlen_n(qq, W)) nn_add_n(mm, mm, qq, W);
	} else if (nn_cmp_n(mm, q2, W) < 0) {
		nn_add_n(mm, mm, q2, W);
		if (nn_bitlen_n(mm, W) == nn_bitlen_n(q2, W)) nn_add_n(mm, mm, q2, W);
	}
	int mlen = nn_bitlen_n(mm, W);
	if (mlen == 0) return -1; /* MUST_HAVE((mlen != 0)) :1623 */
	mlen--;

	pt_t T[3];
	T[0] = *in;
	ret_ops |= pt_add_cf(&T[1], &T[0], &T[0], c);
	while (mlen > 0) {
		--mlen;
		int mbit = nn_getbit_n(mm, mlen);
		pt_t t1;
		ret_ops |= pt_add_cf(&T[2], &T[mbit], &T[mbit], c);
		ret_ops |= pt_add_cf(&t1, &T[0], &T[1], c);
		T[1] = t1;
		T[0] = T[2 - mbit];
		T[1] = T[1 + mbit];
	}
	*out = T[0];
	return ret_ops;
}

/* prj_pt_mul (curves/prj_pt.c:1759-1780): input on-curve check, ladder, output on-curve check */
static int pt_mul(pt_t *out, const u64 *m, int mlimbs, const pt_t *in, const curve_t *c)
{
	if (!pt_is_on_curve(in, c)) return -1;
	if (pt_mul_ladder(out, m, mlimbs, in, c)) return -1;
	if (!pt_is_on_curve(out, c)) return -1;
	return 0;
}

/* prj_pt_unique (curves/prj_pt.c:241-273): X/Z, Y/Z, 1; error on infinity */
static int pt_unique(pt_t *out, const pt_t *in, const curve_t *c)
{
	u64 zi[MAXL];
	if (pt_iszero(in, c)) return -1;
	mod_inv_fermat(zi, in->Z, &c->fp);
	mod_mul(out->Y, in->Y, zi, &c->fp);
	mod_mul(out->X, in->X, zi, &c->fp);
	memset(out->Z, 0, sizeof(out->Z));
	out->Z[0] = 1;
	return 0;
}

/* prj_pt_import_from_aff_buf (curves/prj_pt.c:511-551): x,y < p (fp_import_from_buf) and on curve */
static int pt_import_aff(pt_t *o, const uint8_t *buf, const curve_t *c)
{
	nn_from_be(o->X, c->n, buf, c->plen);
	nn_from_be(o->Y, c->n, buf + c->plen, c->plen);
	memset(o->Z, 0, sizeof(o->Z));
	o->Z[0] = 1;
	if (nn_cmp_n(o->X, c->fp.p, c->n) >= 0 || nn_cmp_n(o->Y, c->fp.p, c->n) >= 0) return -1;
	if (!pt_is_on_curve(o, c)) return -1;
	return 0;
}

/* ------------------------------------------------------------------------------------------------ batch drivers */

typedef struct {
	const curve_t *c;
	uint32_t lo, hi;
	const uint8_t *scalars;
	uint32_t slen;
	const uint8_t *points;
	uint8_t *out;
	int8_t *status;
	/* ecdsa */
	const uint8_t *sigs, *pubkeys, *digests, *privkeys, *nonces;
	uint32_t hlen;
	uint8_t *sigs_out;
	u64 mul_count;
} job_t;

static void *smul_worker(void *arg)
{
	job_t *j = (job_t *)arg;
	const curve_t *c = j->c;
	g_mul_count = 0;
	for (uint32_t i = j->lo; i < j->hi; i++) {
		u64 k[BIGL];
		pt_t in, out, aff;
		uint8_t *o = j->out + (size_t)i * 2 * c->plen;
		memset(o, 0, 2 * c->plen);
		j->status[i] = -1;
		if (j->slen > 8 * BIGL) continue;
		nn_from_be(k, BIGL, j->scalars + (size_t)i * j->slen, j->slen);
		if (j->points) {
			if (pt_import_aff(&in, j->points + (size_t)i * 2 * c->plen, c)) continue;
		} else {
			memcpy(in.X, c->gx, sizeof(in.X));
			memcpy(in.Y, c->gy, sizeof(in.Y));
			memset(in.Z, 0, sizeof(in.Z));
			in.Z[0] = 1;
		}
		g_mul_count = 0;
		if (pt_mul(&out, k, BIGL, &in, c)) continue;
		j->mul_count = g_mul_count;
		if (pt_iszero(&out, c)) {
			j->status[i] = 1;
			continue;
		}
		if (pt_unique(&aff, &out, c)) continue;
		nn_to_be(o, c->plen, aff.X, c->n);
		nn_to_be(o + c->plen, c->plen, aff.Y, c->n);
		j->status[i] = 0;
	}
	return NULL;
}

/* reduce x (n limbs, < 2^(64n)) modulo q by repeated subtraction; all three target curves have 2^(64n) < 2q */
static void reduce_mod_q(u64 *x, const curve_t *c)
{
	while (nn_cmp_n(x, c->fq.p, c->n) >= 0) nn_sub_n(x, x, c->fq.p, c->n);
}

/* steps 3-4 of __ecdsa_verify_finalize (sig/ecdsa_common.c:760-777): e = (OS2I(h) >> max(0, 8*hsize - |q|)) mod q */
static void digest_to_e(u64 *e, const uint8_t *h, uint32_t hlen, const curve_t *c)
{
	u64 big[16];
	int rshift = 0;
	memset(big, 0, sizeof(big));
	nn_from_be(big, 16, h, hlen > 128 ? 128 : hlen);
	if ((int)(hlen * 8) > c->qbits) rshift = (int)(hlen * 8) - c->qbits;
	if (rshift) { /* nn_rshift_fixedlen (nn/nn_logical.c:151) */
		int ws = rshift / 64, bs = rshift % 64;
		for (int i = 0; i < 16; i++) {
			u64 lo = (i + ws < 16) ? big[i + ws] : 0;
			u64 hi = (i + ws + 1 < 16) ? big[i + ws + 1] : 0;
			big[i] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
		}
	}
	memcpy(e, big, sizeof(u64) * (size_t)c->n); /* now < 2^qbits <= 2^(64n) */
	reduce_mod_q(e, c);
}

/* __ecdsa_verify_init checks (sig/ecdsa_common.c:645-658) + __ecdsa_verify_finalize (:702-840) */
static int ecdsa_verify_one(const uint8_t *sig, const uint8_t *pub, const uint8_t *h, uint32_t hlen,
			    const curve_t *c)
{
	u64 r[MAXL], s[MAXL], e[MAXL], sinv[MAXL], u[MAXL], v[MAXL], rp[MAXL];
	pt_t Y, G, uG, vY, W, Wa;
	int n = c->n;
	nn_from_be(r, n, sig, c->qlen);
	nn_from_be(s, n, sig + c->qlen, c->qlen);
	/* 1. reject r or s == 0 or >= q */
	if (nn_iszero_n(r, n) || nn_iszero_n(s, n)) return -1;
	if (nn_cmp_n(r, c->fq.p, n) >= 0 || nn_cmp_n(s, c->fq.p, n) >= 0) return -1;
	/* public key import: ec_pub_key_import_from_aff_buf (sig/ec_key.c:181-214) */
	if (pt_import_aff(&Y, pub, c)) return -1;
	digest_to_e(e, h, hlen, c);
	mod_inv_fermat(sinv, s, &c->fq);       /* sinv = s^-1 mod q      (:781) */
	mod_mul(u, e, sinv, &c->fq);           /* u = e*sinv mod q       (:786) */
	memcpy(G.X, c->gx, sizeof(G.X));
	memcpy(G.Y, c->gy, sizeof(G.Y));
	memset(G.Z, 0, sizeof(G.Z));
	G.Z[0] = 1;
	if (pt_mul(&uG, u, n, &G, c)) return -1; /* (:788) */
	mod_mul(v, r, sinv, &c->fq);           /* v = r*sinv mod q       (:791) */
	if (pt_mul(&vY, v, n, &Y, c)) return -1; /* (:793) */
	if (pt_add_cf(&W, &uG, &vY, c)) return -1; /* (:796) */
	if (pt_iszero(&W, c)) return -1;       /* (:799-800) */
	if (pt_unique(&Wa, &W, c)) return -1;  /* (:803) */
	memcpy(rp, Wa.X, sizeof(rp));
	reduce_mod_q(rp, c);                   /* r' = W'_x mod q        (:806) */
	return nn_cmp_n(rp, r, n) == 0 ? 0 : -1; /* (:809-810) */
}

/* x mod q for a wide x (nlimbs limbs): plain binary long division, the result of nn_mod (nn/nn_div.c:1005) */
static void wide_mod_q(u64 *out, const u64 *x, int nlimbs, const curve_t *c)
{
	u64 r[MAXL + 1];
	int n = c->n;
	memset(r, 0, sizeof(r));
	for (int bit = 64 * nlimbs - 1; bit >= 0; bit--) {
		u64 carry = (x[bit / 64] >> (bit % 64)) & 1;
		for (int i = 0; i <= n; i++) {
			u64 nc = r[i] >> 63;
			r[i] = (r[i] << 1) | carry;
			carry = nc;
		}
		if (r[n] || nn_cmp_n(r, c->fq.p, n) >= 0) {
			u64 bw = nn_sub_n(r, r, c->fq.p, n);
			r[n] -= bw;
		}
	}
	memcpy(out, r, sizeof(u64) * (size_t)n);
}

/* _ecfsdsa_verify_init checks (sig/ecfsdsa.c:447-470) + _ecfsdsa_verify_finalize (:536-610) on h = H(r || m):
 * signature = r || s with r = W_x || W_y (2*plen bytes) and s (qlen bytes). */
static int ecfsdsa_verify_one(const uint8_t *sig, const uint8_t *pub, const uint8_t *h, uint32_t hlen,
			      const curve_t *c)
{
	u64 s[MAXL], e[MAXL], big[16];
	pt_t Y, G, R, sG, eY, W, Wa;
	uint8_t rprime[2 * 8 * MAXL];
	int n = c->n;
	/* 1. r must be a point of the curve (coordinates < p, :453-460) */
	if (pt_import_aff(&R, sig, c)) return -1;
	/* 2. s in ]0, q[ (:465-470) */
	nn_from_be(s, n, sig + 2 * c->plen, c->qlen);
	if (nn_iszero_n(s, n) || nn_cmp_n(s, c->fq.p, n) >= 0) return -1;
	if (pt_import_aff(&Y, pub, c)) return -1;
	/* 4. e = -(OS2I(h) mod q) mod q, the WHOLE digest (:590-594) */
	memset(big, 0, sizeof(big));
	nn_from_be(big, 16, h, hlen > 128 ? 128 : hlen);
	wide_mod_q(e, big, 16, c);
	if (!nn_iszero_n(e, n)) nn_sub_n(e, c->fq.p, e, n);
	/* 5. W' = sG + eY (:597-600) */
	memcpy(G.X, c->gx, sizeof(G.X));
	memcpy(G.Y, c->gy, sizeof(G.Y));
	memset(G.Z, 0, sizeof(G.Z));
	G.Z[0] = 1;
	if (pt_mul(&sG, s, n, &G, c)) return -1;
	if (pt_mul(&eY, e, n, &Y, c)) return -1;
	if (pt_add_cf(&W, &sG, &eY, c)) return -1;
	if (pt_iszero(&W, c)) return -1;       /* prj_pt_unique fails on the point at infinity (curves/prj_pt.c:246) */
	if (pt_unique(&Wa, &W, c)) return -1;
	/* 6.-7. r' = FE2OS(W'_x) || FE2OS(W'_y) must equal r (:603-610) */
	nn_to_be(rprime, c->plen, Wa.X, n);
	nn_to_be(rprime + c->plen, c->plen, Wa.Y, n);
	return memcmp(rprime, sig, 2 * c->plen) == 0 ? 0 : -1;
}

/* _bip0340_verify_init checks (sig/bip0340.c:383-465) + _bip0340_verify_finalize (:497-577) on
 * h = H(H(tag) || H(tag) || r || x(Y) || m), the tagged challenge hash computed by the caller: 0 valid, -1 invalid */
static int bip0340_verify_one(const uint8_t *sig, const uint8_t *pub, const uint8_t *h, uint32_t hlen,
			      const curve_t *c)
{
	u64 r[MAXL], s[MAXL], e[MAXL], big[16];
	pt_t Y, G, sG, eY, W, Wa;
	int n = c->n;
	/* r is imported as a field element: it must be < p (fp_import_from_buf, :431) */
	nn_from_be(r, n, sig, c->plen);
	if (nn_cmp_n(r, c->fp.p, n) >= 0) return -1;
	/* s < q (:433-434); zero is not excluded */
	nn_from_be(s, n, sig + c->plen, c->qlen);
	if (nn_cmp_n(s, c->fq.p, n) >= 0) return -1;
	if (pt_import_aff(&Y, pub, c)) return -1;
	/* e = OS2I(h) mod q with the whole digest (:530-531), then -e mod q (:538) */
	memset(big, 0, sizeof(big));
	nn_from_be(big, 16, h, hlen > 128 ? 128 : hlen);
	wide_mod_q(e, big, 16, c);
	if (!nn_iszero_n(e, n)) nn_sub_n(e, c->fq.p, e, n);
	/* lift the key to its even-y representative (:540-545) */
	if (Y.Y[0] & 1) nn_sub_n(Y.Y, c->fp.p, Y.Y, n);
	memcpy(G.X, c->gx, sizeof(G.X));
	memcpy(G.Y, c->gy, sizeof(G.Y));
	memset(G.Z, 0, sizeof(G.Z));
	G.Z[0] = 1;
	if (pt_mul(&sG, s, n, &G, c)) return -1;
	if (pt_mul(&eY, e, n, &Y, c)) return -1;
	if (pt_add_cf(&W, &sG, &eY, c)) return -1;
	if (pt_iszero(&W, c)) return -1;          /* prj_pt_unique fails on infinity (:552), iszero check (:555-556) */
	if (pt_unique(&Wa, &W, c)) return -1;
	if (Wa.Y[0] & 1) return -1;               /* odd y (:559-560) */
	return nn_cmp_n(Wa.X, r, n) == 0 ? 0 : -1; /* x(W') == r (:563-564) */
}

static void *bip_verify_worker(void *arg)
{
	job_t *j = (job_t *)arg;
	const curve_t *c = j->c;
	for (uint32_t i = j->lo; i < j->hi; i++) {
		j->status[i] = (int8_t)bip0340_verify_one(j->sigs + (size_t)i * (c->plen + c->qlen),
							  j->pubkeys + (size_t)i * 2 * c->plen,
							  j->digests + (size_t)i * j->hlen, j->hlen, c);
	}
	return NULL;
}

/* W = a*G + b*Y as the reference's Schnorr-type verifications compute it (e.g. sig/ecsdsa_common.c:493-497:
 * prj_pt_mul, prj_pt_mul, prj_pt_add, prj_pt_unique): 0 finite (affine bytes written), 1 infinity, -1 key rejected */
static int double_smul_one(uint8_t *out, const uint8_t *ab, const uint8_t *pub, const curve_t *c)
{
	u64 a[MAXL], b[MAXL];
	pt_t Y, G, aG, bY, W, Wa;
	int n = c->n;
	memset(out, 0, 2 * c->plen);
	nn_from_be(a, n, ab, c->qlen);
	nn_from_be(b, n, ab + c->qlen, c->qlen);
	if (pt_import_aff(&Y, pub, c)) return -1;
	memcpy(G.X, c->gx, sizeof(G.X));
	memcpy(G.Y, c->gy, sizeof(G.Y));
	memset(G.Z, 0, sizeof(G.Z));
	G.Z[0] = 1;
	if (pt_mul(&aG, a, n, &G, c)) return -1;
	if (pt_mul(&bY, b, n, &Y, c)) return -1;
	if (pt_add_cf(&W, &aG, &bY, c)) return -1;
	if (pt_iszero(&W, c)) return 1;
	if (pt_unique(&Wa, &W, c)) return -1;
	nn_to_be(out, c->plen, Wa.X, n);
	nn_to_be(out + c->plen, c->plen, Wa.Y, n);
	return 0;
}

static void *double_smul_worker(void *arg)
{
	job_t *j = (job_t *)arg;
	const curve_t *c = j->c;
	for (uint32_t i = j->lo; i < j->hi; i++)
		j->status[i] = (int8_t)double_smul_one(j->out + (size_t)i * 2 * c->plen, j->sigs + (size_t)i * 2 * c->qlen,
						       j->pubkeys + (size_t)i * 2 * c->plen, c);
	return NULL;
}

static void *fs_verify_worker(void *arg)
{
	job_t *j = (job_t *)arg;
	const curve_t *c = j->c;
	for (uint32_t i = j->lo; i < j->hi; i++) {
		j->status[i] = (int8_t)ecfsdsa_verify_one(j->sigs + (size_t)i * (2 * c->plen + c->qlen),
							  j->pubkeys + (size_t)i * 2 * c->plen,
							  j->digests + (size_t)i * j->hlen, j->hlen, c);
	}
	return NULL;
}

static void *verify_worker(void *arg)
{
	job_t *j = (job_t *)arg;
	const curve_t *c = j->c;
	for (uint32_t i = j->lo; i < j->hi; i++) {
		j->status[i] = (int8_t)ecdsa_verify_one(j->sigs + (size_t)i * 2 * c->qlen,
							j->pubkeys + (size_t)i * 2 * c->plen,
							j->digests + (size_t)i * j->hlen, j->hlen, c);
	}
	return NULL;
}

/* __ecdsa_sign_finalize (sig/ecdsa_common.c:318-586) with the nonce supplied by the caller (the reference's
 * test harness injects it the same way, tests/ec_self_tests_core.h:34): r = x(kG) mod q, s = k^-1 (x r + e) mod q.
 * Returns 0, -1 (d or k out of range) or 2 where the reference would "goto restart" with a new nonce. */
static int ecdsa_sign_one(uint8_t *sig, const uint8_t *priv, const uint8_t *nonce, const uint8_t *h,
			  uint32_t hlen, const curve_t *c)
{
	u64 d[MAXL], k[MAXL], e[MAXL], r[MAXL], s[MAXL], kinv[MAXL], t[MAXL];
	pt_t G, kG, A;
	int n = c->n;
	nn_from_be(d, n, priv, c->qlen);
	nn_from_be(k, n, nonce, c->qlen);
	if (nn_iszero_n(d, n) || nn_cmp_n(d, c->fq.p, n) >= 0) return -1;
	if (nn_iszero_n(k, n) || nn_cmp_n(k, c->fq.p, n) >= 0) return -1;
	digest_to_e(e, h, hlen, c);
	memcpy(G.X, c->gx, sizeof(G.X));
	memcpy(G.Y, c->gy, sizeof(G.Y));
	memset(G.Z, 0, sizeof(G.Z));
	G.Z[0] = 1;
	if (pt_mul(&kG, k, n, &G, c)) return -1;
	if (pt_unique(&A, &kG, c)) return -1;
	memcpy(r, A.X, sizeof(r));
	reduce_mod_q(r, c);
	if (nn_iszero_n(r, n)) return 2;               /* 7. r == 0 -> restart (:487) */
	mod_mul(t, d, r, &c->fq);
	if (nn_cmp_n(e, t, n) == 0) return 2;          /* 8. e == rx -> restart (:513) */
	mod_add(t, t, e, &c->fq);
	mod_inv_fermat(kinv, k, &c->fq);
	mod_mul(s, kinv, t, &c->fq);
	if (nn_iszero_n(s, n)) return 2;               /* 10. s == 0 -> restart (:545) */
	nn_to_be(sig, c->qlen, r, n);
	nn_to_be(sig + c->qlen, c->qlen, s, n);
	return 0;
}

static void *sign_worker(void *arg)
{
	job_t *j = (job_t *)arg;
	const curve_t *c = j->c;
	for (uint32_t i = j->lo; i < j->hi; i++) {
		j->status[i] = (int8_t)ecdsa_sign_one(j->sigs_out + (size_t)i * 2 * c->qlen,
						      j->privkeys + (size_t)i * c->qlen, j->nonces + (size_t)i * c->qlen,
						      j->digests + (size_t)i * j->hlen, j->hlen, c);
	}
	return NULL;
}

static void run_jobs(void *(*fn)(void *), job_t *proto, uint32_t n, int nthreads)
{
	if (nthreads < 1) nthreads = 1;
	if ((uint32_t)nthreads > n) nthreads = n ? (int)n : 1;
	pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
	job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
	for (int t = 0; t < nthreads; t++) {
		jobs[t] = *proto;
		jobs[t].lo = (uint32_t)(((uint64_t)n * (uint64_t)t) / (uint64_t)nthreads);
		jobs[t].hi = (uint32_t)(((uint64_t)n * (uint64_t)(t + 1)) / (uint64_t)nthreads);
		pthread_create(&th[t], NULL, fn, &jobs[t]);
	}
	for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
	proto->mul_count = jobs[0].mul_count;
	free(th);
	free(jobs);
}

int ora_prj_pt_mul_batch(const char *curve, uint32_t n, const uint8_t *scalars, uint32_t slen,
			 const uint8_t *points, uint8_t *out, int8_t *status, int nthreads)
{
	curve_t c;
	job_t p;
	if (curve_load(&c, curve)) return -1;
	memset(&p, 0, sizeof(p));
	p.c = &c;
	p.scalars = scalars;
	p.slen = slen;
	p.points = points;
	p.out = out;
	p.status = status;
	run_jobs(smul_worker, &p, n, nthreads);
	g_last_mul_count = p.mul_count;
	return 0;
}

uint64_t ora_last_mul_count(void) { return g_last_mul_count; }

int ora_ecdsa_verify_digest_batch(const char *curve, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				  const uint8_t *digests, uint32_t hlen, int8_t *verdict, int nthreads)
{
	curve_t c;
	job_t p;
	if (curve_load(&c, curve)) return -1;
	memset(&p, 0, sizeof(p));
	p.c = &c;
	p.sigs = sigs;
	p.pubkeys = pubkeys;
	p.digests = digests;
	p.hlen = hlen;
	p.status = verdict;
	run_jobs(verify_worker, &p, n, nthreads);
	return 0;
}

/* ECFSDSA verification on h = H(r || m) (sig/ecfsdsa.c); sigs are [n][2*plen + qlen] */
int ora_ecfsdsa_verify_digest_batch(const char *curve, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				  const uint8_t *digests, uint32_t hlen, int8_t *verdict, int nthreads)
{
	curve_t c;
	job_t p;
	if (curve_load(&c, curve)) return -1;
	memset(&p, 0, sizeof(p));
	p.c = &c;
	p.sigs = sigs;
	p.pubkeys = pubkeys;
	p.digests = digests;
	p.hlen = hlen;
	p.status = verdict;
	run_jobs(fs_verify_worker, &p, n, nthreads);
	return 0;
}

/* W_i = a_i*G + b_i*Y_i, affine (ab: [n][2*qlen]) */
int ora_double_smul_batch(const char *curve, uint32_t n, const uint8_t *ab, const uint8_t *pubkeys, uint8_t *out,
			  int8_t *status, int nthreads)
{
	curve_t c;
	job_t p;
	if (curve_load(&c, curve)) return -1;
	memset(&p, 0, sizeof(p));
	p.c = &c;
	p.sigs = ab;
	p.pubkeys = pubkeys;
	p.out = out;
	p.status = status;
	run_jobs(double_smul_worker, &p, n, nthreads);
	return 0;
}

/* BIP0340 verification on the tagged challenge hash (sig/bip0340.c); sigs are [n][plen + qlen] (r, s) */
int ora_bip0340_verify_digest_batch(const char *curve, uint32_t n, const uint8_t *sigs, const uint8_t *pubkeys,
				    const uint8_t *digests, uint32_t hlen, int8_t *verdict, int nthreads)
{
	curve_t c;
	job_t p;
	if (curve_load(&c, curve)) return -1;
	memset(&p, 0, sizeof(p));
	p.c = &c;
	p.sigs = sigs;
	p.pubkeys = pubkeys;
	p.digests = digests;
	p.hlen = hlen;
	p.status = verdict;
	run_jobs(bip_verify_worker, &p, n, nthreads);
	return 0;
}

int ora_ecdsa_sign_digest_batch(const char *curve, uint32_t n, const uint8_t *privkeys, const uint8_t *nonces,
				const uint8_t *digests, uint32_t hlen, uint8_t *sigs, int8_t *status, int nthreads)
{
	curve_t c;
	job_t p;
	if (curve_load(&c, curve)) return -1;
	memset(&p, 0, sizeof(p));
	p.c = &c;
	p.privkeys = privkeys;
	p.nonces = nonces;
	p.digests = digests;
	p.hlen = hlen;
	p.sigs_out = sigs;
	p.status = status;
	run_jobs(sign_worker, &p, n, nthreads);
	return 0;
}

int ora_fp_mul_monty(const char *curve, const uint8_t *a, const uint8_t *b, uint8_t *out)
{
	curve_t c;
	u64 x[MAXL], y[MAXL], z[MAXL];
	if (curve_load(&c, curve)) return -1;
	nn_from_be(x, c.n, a, c.plen);
	nn_from_be(y, c.n, b, c.plen);
	if (nn_cmp_n(x, c.fp.p, c.n) >= 0 || nn_cmp_n(y, c.fp.p, c.n) >= 0) return -1;
	mul_redc1(z, x, y, &c.fp);
	nn_to_be(out, c.plen, z, c.n);
	return 0;
}
