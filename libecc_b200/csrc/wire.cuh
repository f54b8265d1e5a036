/*
 * wire.cuh — the reference's *structured* key / signature records on the device (the on-disk / wire
 * step on either side of the scalar-multiplication path).
 *
 * Record layouts (paths relative to /root/reference/src):
 *   structured public key   [EC_PUBKEY = 0][ec_alg_type][ec_curve_type] X || Y || Z   (3 + 3*plen bytes; homogeneous
 *                           projective, big-endian)          ec_structured_pub_key_export_to_buf  sig/ec_key.c:451
 *   structured private key  [EC_PRIVKEY = 1][ec_alg_type][ec_curve_type] x            (3 + L bytes, big-endian, L =
 *                           EC_PRIV_KEY_EXPORT_SIZE, sig/ec_key.h:75-83)  ec_structured_priv_key_export_to_buf :358
 *   structured signature    [ec_alg_type][hash_alg_type][ec_curve_type] r || s        (3 + 2*qlen bytes)
 *                                                            ec_structured_sig_export_to_buf      sig/sig_algs.c:742
 * The 3-byte header makes every field of a record array unaligned, so the records are first unpacked into the packed,
 * aligned field arrays the arithmetic kernels read (one thread per payload byte: fully coalesced, ~100-200 B per
 * item against the ~0.5 M integer multiply-adds of the operation that follows).
 */
#pragma once
#include <stdint.h>

namespace eccb200 {

/* out[i][0..out_len) = the LAST out_len bytes of the payload of record i (payload = bytes [3, 3 + payload_len)).
 * payload_len >= out_len; the bytes in front (a private key exported on more bytes than qlen) are checked by
 * k_struct_check. */
__global__ void k_struct_unpack(uint32_t n, const uint8_t *__restrict__ rec, uint32_t stride, uint32_t payload_len,
				uint8_t *__restrict__ out, uint32_t out_len)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (uint64_t)n * out_len) return;
	const uint32_t i = (uint32_t)(t / out_len), b = (uint32_t)(t % out_len);
	out[t] = rec[(size_t)i * stride + 3 + (payload_len - out_len) + b];
}

/* state[i] = -1 when the header of record i is not (h0, h1, h2) — the checks of ec_structured_pub_key_import_from_buf
 * (sig/ec_key.c:429-441) / ..._priv_key_import_from_buf (:326-338) — or when one of the `lead` payload bytes in front
 * of the unpacked field is non-zero (then the integer is >= 2^(8*qlen) > q).  Other entries are left untouched. */
__global__ void k_struct_check(uint32_t n, const uint8_t *__restrict__ rec, uint32_t stride, uint8_t h0, uint8_t h1,
			       uint8_t h2, uint32_t lead, int8_t *__restrict__ state)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint8_t *r = rec + (size_t)i * stride;
	bool ok = r[0] == h0 && r[1] == h1 && r[2] == h2;
	for (uint32_t b = 0; b < lead; b++) ok = ok && r[3 + b] == 0;
	if (!ok) state[i] = -1;
}

/* Structured public key records from affine points: header, then X || Y || Z with Z = 1; the point at infinity
 * (state 1) is written as (0, 1, 0) (prj_pt_zero, curves/prj_pt.c:124-136) and a rejected slot (state < 0) as an
 * all-zero payload.  One thread per record byte. */
__global__ void k_struct_pack_pub(uint32_t n, const uint8_t *__restrict__ aff, uint32_t plen,
				  const int8_t *__restrict__ state, uint8_t h0, uint8_t h1, uint8_t h2,
				  uint8_t *__restrict__ rec)
{
	const uint32_t stride = 3 + 3 * plen;
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (uint64_t)n * stride) return;
	const uint32_t i = (uint32_t)(t / stride), b = (uint32_t)(t % stride);
	const int st = state[i];
	uint8_t v;
	if (b < 3) {
		v = b == 0 ? h0 : (b == 1 ? h1 : h2);
	} else {
		const uint32_t o = b - 3;
		if (st < 0) v = 0;
		else if (st == 1) v = (o == 2 * plen - 1) ? 1 : 0;                   /* (0, 1, 0) */
		else if (o < 2 * plen) v = aff[(size_t)i * 2 * plen + o];
		else v = (o == 3 * plen - 1) ? 1 : 0;                                 /* Z = 1 */
	}
	rec[t] = v;
}

/* Generic record writer: header + payload (a signature r || s); payload zeroed when state[i] != 0. */
__global__ void k_struct_pack(uint32_t n, const uint8_t *__restrict__ payload, uint32_t payload_len,
			      const int8_t *__restrict__ state, uint8_t h0, uint8_t h1, uint8_t h2,
			      uint8_t *__restrict__ rec)
{
	const uint32_t stride = 3 + payload_len;
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (uint64_t)n * stride) return;
	const uint32_t i = (uint32_t)(t / stride), b = (uint32_t)(t % stride);
	uint8_t v;
	if (b < 3) v = b == 0 ? h0 : (b == 1 ? h1 : h2);
	else v = state[i] == 0 ? payload[(size_t)i * payload_len + (b - 3)] : 0;
	rec[t] = v;
}

} // namespace eccb200
