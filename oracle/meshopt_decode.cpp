/* meshopt_decode.cpp — CPU restatement of the EXT_meshopt_compression decoders the reference runs at load time
 * (src/vk_gltf_viewer/assets.cpp:111-171: meshopt_decodeVertexBuffer / IndexBuffer / IndexSequence + the oct / quat / exp
 * filters), from meshoptimizer @ the reference's pinned submodule:
 *   vertex codec   submodules/meshoptimizer/src/vertexcodec.cpp:300-415 (byte groups, blocks), :1178-1240 (stream framing)
 *   index codec    submodules/meshoptimizer/src/indexcodec.cpp:362-540 (triangle list), :618-672 (sequence), :95-135 (varints)
 *   filters        submodules/meshoptimizer/src/vertexfilter.cpp:75-160 (the portable scalar definitions)
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY PINNED: tests/test_meshopt_codec.py checks this file against the
 * reference's own known-answer vectors (submodules/meshoptimizer/demo/tests.cpp:24-59,515-629, frozen in
 * tests/golden/meshopt_codec.npz by tests/golden/make_meshopt_golden.py) and against streams encoded AND decoded by the
 * reference's meshoptimizer built from source (oracle/_ref/libmeshopt_ref.so, libmeshopt_ref_nosimd.so).
 *
 * Filters: the reference has one scalar and three SIMD definitions per filter; they differ in the float->int rounding
 * (scalar: x*s +- 0.5 truncated; SSE: cvtps2dq, round-half-even) and, for 8-bit octahedral, in 1/sqrt (SSE: rsqrtps, an
 * approximation whose bits differ between CPU vendors).  This restatement — and the CUDA kernels checked against it —
 * follows the SCALAR definitions (what -DMESHOPTIMIZER_NO_SIMD builds); every known-answer vector of demo/tests.cpp holds
 * for it, and the SSE build differs from it by at most one unit in the last place of a component.
 *
 * Return codes are meshoptimizer's: 0 ok, -1 bad header / version, -2 truncated stream, -3 trailing bytes.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace {

/* ---- vertex codec (vertexcodec.cpp:106-126) ---- */
const unsigned char kVertexHeader = 0xa0;
const size_t kVertexBlockSizeBytes = 8192, kVertexBlockMaxSize = 256, kByteGroupSize = 16, kByteGroupDecodeLimit = 24, kTailMaxSize = 32;

size_t vertex_block_size(size_t vertex_size) { /* :116-126 */
	size_t r = (kVertexBlockSizeBytes / vertex_size) & ~(kByteGroupSize - 1);
	return r < kVertexBlockMaxSize ? r : kVertexBlockMaxSize;
}

/* one 16-byte group (:300-346): bitslog2 0 -> zeros; 1 / 2 -> 2- / 4-bit codes, MSB first, the all-ones code escapes to the
 * next byte of the group's tail; 3 -> 16 literal bytes.  Returns the position after the group. */
const unsigned char* decode_group(const unsigned char* data, unsigned char* out, int bitslog2) {
	if (bitslog2 == 0) { memset(out, 0, kByteGroupSize); return data; }
	if (bitslog2 == 3) { memcpy(out, data, kByteGroupSize); return data + kByteGroupSize; }
	const int bits = bitslog2 == 1 ? 2 : 4, per_byte = 8 / bits, sentinel = (1 << bits) - 1;
	const unsigned char* var = data + kByteGroupSize * bits / 8;
	for (size_t i = 0; i < kByteGroupSize; ++i) {
		const unsigned char byte = data[i / per_byte];
		const int enc = (byte >> (8 - bits * (int)(i % per_byte + 1))) & sentinel;
		out[i] = enc == sentinel ? *var++ : (unsigned char)enc;
	}
	return var;
}

/* one byte plane of a block (:349-376) */
const unsigned char* decode_bytes(const unsigned char* data, const unsigned char* end, unsigned char* out, size_t n) {
	const unsigned char* header = data;
	const size_t header_size = (n / kByteGroupSize + 3) / 4;
	if ((size_t)(end - data) < header_size) return NULL;
	data += header_size;
	for (size_t i = 0; i < n; i += kByteGroupSize) {
		if ((size_t)(end - data) < kByteGroupDecodeLimit) return NULL;
		const size_t g = i / kByteGroupSize;
		data = decode_group(data, out + i, (header[g / 4] >> ((g % 4) * 2)) & 3);
	}
	return data;
}

inline unsigned char unzigzag8(unsigned char v) { return (unsigned char)(-(v & 1) ^ (v >> 1)); } /* :128-131 */

} // namespace

extern "C" {

int orc_meshopt_decode_vertex(void* destination, size_t vertex_count, size_t vertex_size, const unsigned char* buffer, size_t buffer_size) {
	/* :1178-1240 */
	if (vertex_size == 0 || vertex_size > 256 || vertex_size % 4) return -1;
	unsigned char* dst = (unsigned char*)destination;
	const unsigned char* data = buffer;
	const unsigned char* end = buffer + buffer_size;
	if ((size_t)(end - data) < 1 + vertex_size) return -2;
	const unsigned char h = *data++;
	if ((h & 0xf0) != kVertexHeader) return -1;
	if ((h & 0x0f) > 0) return -1;
	unsigned char last[256];
	memcpy(last, end - vertex_size, vertex_size);
	const size_t block = vertex_block_size(vertex_size);
	unsigned char plane[kVertexBlockMaxSize];
	for (size_t off = 0; off < vertex_count;) {
		const size_t n = off + block < vertex_count ? block : vertex_count - off;
		const size_t aligned = (n + kByteGroupSize - 1) & ~(kByteGroupSize - 1);
		for (size_t k = 0; k < vertex_size; ++k) { /* :378-413 */
			data = decode_bytes(data, end, plane, aligned);
			if (!data) return -2;
			unsigned char p = last[k];
			for (size_t i = 0; i < n; ++i) {
				p = (unsigned char)(unzigzag8(plane[i]) + p);
				dst[(off + i) * vertex_size + k] = p;
			}
			last[k] = p;
		}
		off += n;
	}
	const size_t tail = vertex_size < kTailMaxSize ? kTailMaxSize : vertex_size;
	return (size_t)(end - data) == tail ? 0 : -3;
}

/* ---- index codecs ---- */
static unsigned decode_vbyte(const unsigned char*& data) { /* indexcodec.cpp:95-120 */
	unsigned char lead = *data++;
	if (lead < 128) return lead;
	unsigned result = lead & 127, shift = 7;
	for (int i = 0; i < 4; ++i) {
		unsigned char g = *data++;
		result |= (unsigned)(g & 127) << shift;
		shift += 7;
		if (g < 128) break;
	}
	return result;
}
static unsigned decode_index(const unsigned char*& data, unsigned last) { /* :129-135 */
	unsigned v = decode_vbyte(data);
	return last + ((v >> 1) ^ (unsigned)-(int)(v & 1));
}
static void put(void* dst, size_t i, size_t index_size, unsigned v) {
	if (index_size == 2) ((unsigned short*)dst)[i] = (unsigned short)v;
	else ((unsigned*)dst)[i] = v;
}

int orc_meshopt_decode_index(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size) {
	/* :362-540.  State: a 16-entry edge FIFO, a 16-entry vertex FIFO, `next` (the next never-seen index), `last` (delta base). */
	if (index_count % 3 || (index_size != 2 && index_size != 4)) return -1;
	if (buffer_size < 1 + index_count / 3 + 16) return -2;
	if ((buffer[0] & 0xf0) != 0xe0) return -1;
	const int version = buffer[0] & 0x0f;
	if (version > 1) return -1;
	unsigned ef[16][2], vf[16];
	memset(ef, -1, sizeof(ef));
	memset(vf, -1, sizeof(vf));
	size_t eo = 0, vo = 0;
	unsigned next = 0, last = 0;
	const int fecmax = version >= 1 ? 13 : 15;
	const unsigned char* code = buffer + 1;
	const unsigned char* data = code + index_count / 3;
	const unsigned char* safe_end = buffer + buffer_size - 16;
	const unsigned char* aux = safe_end;
	auto push_v = [&](unsigned v, int cond) { vf[vo] = v; vo = (vo + cond) & 15; };
	auto push_e = [&](unsigned a, unsigned b) { ef[eo][0] = a; ef[eo][1] = b; eo = (eo + 1) & 15; };
	for (size_t i = 0; i < index_count; i += 3) {
		if (data > safe_end) return -2;
		const unsigned char ct = *code++;
		unsigned a, b, c;
		if (ct < 0xf0) { /* an edge of the FIFO + one vertex */
			const int fe = ct >> 4, fec = ct & 15;
			a = ef[(eo - 1 - fe) & 15][0];
			b = ef[(eo - 1 - fe) & 15][1];
			if (fec < fecmax) {
				const unsigned cf = vf[(vo - 1 - fec) & 15];
				c = fec == 0 ? next : cf;
				const int fec0 = fec == 0;
				next += fec0;
				push_v(c, fec0);
			} else {
				last = c = fec != 15 ? last + (unsigned)(fec - (fec ^ 3)) : decode_index(data, last);
				push_v(c, 1);
			}
			push_e(c, b);
			push_e(a, c);
		} else if (ct < 0xfe) { /* a new vertex + two coded through the 16-entry table at the end of the stream */
			const unsigned char ca = aux[ct & 15];
			const int feb = ca >> 4, fec = ca & 15;
			a = next++;
			const unsigned bf = vf[(vo - feb) & 15];
			b = feb == 0 ? next : bf;
			const int feb0 = feb == 0;
			next += feb0;
			const unsigned cf = vf[(vo - fec) & 15];
			c = fec == 0 ? next : cf;
			const int fec0 = fec == 0;
			next += fec0;
			push_v(a, 1); push_v(b, feb0); push_v(c, fec0);
			push_e(b, a); push_e(c, b); push_e(a, c);
		} else { /* explicit aux byte; 0 resets `next` */
			const unsigned char ca = *data++;
			const int fea = ct == 0xfe ? 0 : 15, feb = ca >> 4, fec = ca & 15;
			if (ca == 0) next = 0;
			a = fea == 0 ? next++ : 0;
			b = feb == 0 ? next++ : vf[(vo - feb) & 15];
			c = fec == 0 ? next++ : vf[(vo - fec) & 15];
			if (fea == 15) last = a = decode_index(data, last);
			if (feb == 15) last = b = decode_index(data, last);
			if (fec == 15) last = c = decode_index(data, last);
			push_v(a, 1); push_v(b, (feb == 0) | (feb == 15)); push_v(c, (fec == 0) | (fec == 15));
			push_e(b, a); push_e(c, b); push_e(a, c);
		}
		put(destination, i, index_size, a); put(destination, i + 1, index_size, b); put(destination, i + 2, index_size, c);
	}
	return data == safe_end ? 0 : -3;
}

int orc_meshopt_decode_sequence(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size) {
	/* :618-672: varint (delta zig-zag << 1 | baseline) against one of two running baselines */
	if (index_size != 2 && index_size != 4) return -1;
	if (buffer_size < 1 + index_count + 4) return -2;
	if ((buffer[0] & 0xf0) != 0xd0) return -1;
	if ((buffer[0] & 0x0f) > 1) return -1;
	const unsigned char* data = buffer + 1;
	const unsigned char* safe_end = buffer + buffer_size - 4;
	unsigned last[2] = {0, 0};
	for (size_t i = 0; i < index_count; ++i) {
		if (data >= safe_end) return -2;
		unsigned v = decode_vbyte(data);
		const unsigned cur = v & 1;
		v >>= 1;
		const unsigned index = last[cur] + ((v >> 1) ^ (unsigned)-(int)(v & 1));
		last[cur] = index;
		put(destination, i, index_size, index);
	}
	return data == safe_end ? 0 : -3;
}

/* ---- filters, scalar definitions (vertexfilter.cpp:75-160) ---- */
void orc_meshopt_filter_oct(void* buffer, size_t count, size_t stride) { /* stride 4: int8 x4, stride 8: int16 x4 */
	for (size_t i = 0; i < count; ++i) {
		float x, y, zc, mx;
		if (stride == 4) { const signed char* d = (const signed char*)buffer + i * 4; x = d[0]; y = d[1]; zc = d[2]; mx = 127.f; }
		else { const short* d = (const short*)buffer + i * 4; x = d[0]; y = d[1]; zc = d[2]; mx = 32767.f; }
		float z = zc - fabsf(x) - fabsf(y);
		const float t = (z >= 0.f) ? 0.f : z;
		x += (x >= 0.f) ? t : -t;
		y += (y >= 0.f) ? t : -t;
		const float l = sqrtf(x * x + y * y + z * z);
		const float s = mx / l;
		const int xf = (int)(x * s + (x >= 0.f ? 0.5f : -0.5f));
		const int yf = (int)(y * s + (y >= 0.f ? 0.5f : -0.5f));
		const int zf = (int)(z * s + (z >= 0.f ? 0.5f : -0.5f));
		if (stride == 4) { signed char* d = (signed char*)buffer + i * 4; d[0] = (signed char)xf; d[1] = (signed char)yf; d[2] = (signed char)zf; }
		else { short* d = (short*)buffer + i * 4; d[0] = (short)xf; d[1] = (short)yf; d[2] = (short)zf; }
	}
}

void orc_meshopt_filter_quat(void* buffer, size_t count) { /* int16 x4 */
	const float scale = 1.f / sqrtf(2.f);
	short* data = (short*)buffer;
	for (size_t i = 0; i < count; ++i) {
		const int sf = data[i * 4 + 3] | 3;
		const float ss = scale / (float)sf;
		const float x = (float)data[i * 4 + 0] * ss, y = (float)data[i * 4 + 1] * ss, z = (float)data[i * 4 + 2] * ss;
		const float ww = 1.f - x * x - y * y - z * z;
		const float w = sqrtf(ww >= 0.f ? ww : 0.f);
		const int xf = (int)(x * 32767.f + (x >= 0.f ? 0.5f : -0.5f));
		const int yf = (int)(y * 32767.f + (y >= 0.f ? 0.5f : -0.5f));
		const int zf = (int)(z * 32767.f + (z >= 0.f ? 0.5f : -0.5f));
		const int wf = (int)(w * 32767.f + 0.5f);
		const int qc = data[i * 4 + 3] & 3;
		data[i * 4 + ((qc + 1) & 3)] = (short)xf;
		data[i * 4 + ((qc + 2) & 3)] = (short)yf;
		data[i * 4 + ((qc + 3) & 3)] = (short)zf;
		data[i * 4 + ((qc + 0) & 3)] = (short)wf;
	}
}

void orc_meshopt_filter_exp(void* buffer, size_t count_words) { /* 24-bit mantissa, 8-bit exponent -> fp32 */
	unsigned* data = (unsigned*)buffer;
	for (size_t i = 0; i < count_words; ++i) {
		const unsigned v = data[i];
		const int m = (int)(v << 8) >> 8, e = (int)v >> 24;
		union { float f; unsigned ui; } u;
		u.ui = (unsigned)(e + 127) << 23;
		u.f = u.f * (float)m;
		data[i] = u.ui;
	}
}

} // extern "C"
