/*
 * tests/hostsim/modq_hook.cpp — TEST-ONLY: compiles the drop-in's translation unit and exposes its internal mod-q
 * arithmetic (struct ModQ, digest_truncated_mod_q of libecc_b200/csrc/dropin.cpp) to the CPU tests, which compare every
 * operation with Python integers on all eleven curves.  Linked against the engine stub (the drop-in's eccb200_* references
 * must resolve; none is called here).
 */
#include "../../libecc_b200/csrc/dropin.cpp"

static const CurveInfo *curve_by_id(int id)
{
	for (int c = 0; c < kNumCurves; c++)
		if (curves()[c].id == id) return &curves()[c];
	return nullptr;
}

static void be_to_words(uint64_t *o, const uint8_t *be, size_t len)
{
	memset(o, 0, 9 * 8);
	for (size_t j = 0; j < len && j < 72; j++) o[j / 8] |= (uint64_t)be[len - 1 - j] << (8 * (j % 8));
}

extern "C" {

/* op: 0 add, 1 neg(a), 2 a*b mod q (through the Montgomery domain), 3 a^-1 (inv_many on [a, b] -> out = a^-1 || b^-1),
 * 4 to_mont(a), 5 from_mont(a).  a, b, out: qlen big-endian bytes, reduced.  Returns -1 on an unknown curve. */
int modq_op(int curve_id, int op, const uint8_t *a, const uint8_t *b, uint8_t *out)
{
	const CurveInfo *ci = curve_by_id(curve_id);
	if (!ci) return -1;
	const ModQ mq(ci);
	uint64_t x[9], y[9], r[9];
	mq.from_be(x, a);
	mq.from_be(y, b);
	switch (op) {
	case 0: mq.add(r, x, y); break;
	case 1: mq.neg(r, x); break;
	case 2:
		mq.to_mont(x, x);
		mq.mul(r, x, y);
		break;
	case 3: {
		uint64_t v[18];
		memcpy(v, x, sizeof(x));
		memcpy(v + 9, y, sizeof(y));
		mq.inv_many(v, 2);
		mq.to_be(out, v);
		mq.to_be(out + ci->qlen, v + 9);
		return 0;
	}
	case 4: mq.to_mont(r, x); break;
	case 5: mq.from_mont(r, x); break;
	default: return -1;
	}
	mq.to_be(out, r);
	return 0;
}

/* k values (qlen bytes each, nonzero, reduced) inverted together: the chunk-wide inversion of the ECGDSA / ECRDSA adapters */
int modq_inv_many(int curve_id, const uint8_t *vals, uint32_t k, uint8_t *out)
{
	const CurveInfo *ci = curve_by_id(curve_id);
	if (!ci) return -1;
	const ModQ mq(ci);
	std::vector<uint64_t> v((size_t)k * 9);
	for (uint32_t i = 0; i < k; i++) mq.from_be(&v[(size_t)i * 9], vals + (size_t)i * ci->qlen);
	mq.inv_many(v.data(), k);
	for (uint32_t i = 0; i < k; i++) mq.to_be(out + (size_t)i * ci->qlen, &v[(size_t)i * 9]);
	return 0;
}

/* any byte string reduced mod q (nn_init_from_buf + nn_mod) */
int modq_reduce(int curve_id, const uint8_t *be, uint32_t len, uint8_t *out)
{
	const CurveInfo *ci = curve_by_id(curve_id);
	if (!ci) return -1;
	const ModQ mq(ci);
	uint64_t r[9];
	mq.from_be_mod(r, be, len);
	mq.to_be(out, r);
	return 0;
}

/* the ECDSA / ECGDSA digest truncation: leftmost min(8 hlen, bitlen(q)) bits, mod q */
int modq_digest_truncated(int curve_id, const uint8_t *h, uint32_t hlen, uint8_t *out)
{
	const CurveInfo *ci = curve_by_id(curve_id);
	if (!ci || hlen > 128) return -1;
	const ModQ mq(ci);
	DsBatch b;
	b.ci = ci;
	b.hlen = hlen;
	b.qlen = (size_t)ci->qlen;
	b.mq = &mq;
	uint64_t e[9];
	digest_truncated_mod_q(b, e, h, order_bits(ci));
	mq.to_be(out, e);
	return 0;
}

int modq_order_bits(int curve_id)
{
	const CurveInfo *ci = curve_by_id(curve_id);
	return ci ? order_bits(ci) : -1;
}

} /* extern "C" */
