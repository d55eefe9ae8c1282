// meshopt.cu — EXT_meshopt_compression decode on the device.
// Replaces CompressedBufferDataAdapter::ExecuteRange (src/vk_gltf_viewer/assets.cpp:111-171), which runs
// meshopt_decodeVertexBuffer / meshopt_decodeIndexBuffer / meshopt_decodeIndexSequence and the oct / quat / exp filters on the
// host, one enkiTS task per buffer view.  Formats: meshoptimizer @ the reference's pinned submodule —
// vertexcodec.cpp:106-126,300-415,1178-1240, indexcodec.cpp:95-135,362-540,618-672, vertexfilter.cpp:75-160 (scalar definitions).
//
// Vertex streams (the bulk of the bytes).  A stream is a chain of blocks of <= 256 vertices; a block is `stride` byte planes;
// a plane is <= 16 groups of 16 deltas whose encoded size (0 / 4+n / 8+n / 16 bytes) depends on the group's own bytes, so the
// POSITION of everything is only known by walking the stream front to back.  What is serial is therefore split from what is not:
//   vscan    one warp per stream records where every plane starts (u32 per plane): the lanes tabulate the sentinel counts of
//            a 1 KB window for every byte offset, lane 0 walks the groups with one table lookup each — the only serial step;
//   vdecode  one thread block per vertex block: `stride` threads re-walk their plane's <= 16 groups in parallel, then every
//            thread decodes its vertex's byte of four planes at a time (sentinel rank by popcount), a block-wide inclusive scan
//            with byte-wise SIMD adds (__vadd4) undoes the delta coding RELATIVE to the block start, the 8 KB tile is written
//            out with coalesced 4-byte stores, and the block's per-plane total is kept;
//   vcarry   per stream and 4 planes: running sum of the block totals, seeded with the stream's tail vertex;
//   vadd     adds a block's carry to its vertices (byte-wise), completing the chain across blocks.
// Index streams are strictly sequential state machines (two 16-entry FIFOs + two counters): one thread per stream.
// Filters are element-wise kernels over the decoded buffer.
#include "kernels.cuh"

namespace {

__device__ __forceinline__ uint32_t round16(uint32_t n) { return (n + 15u) & ~15u; }
__device__ __forceinline__ uint32_t unzigzag8x(uint32_t v) { return (0u - (v & 1u)) ^ (v >> 1); } // low 8 bits are the result

// encoded size of one 16-delta group (vertexcodec.cpp:300-346): the all-ones codes escape to one tail byte each
__device__ __forceinline__ uint32_t group_size(const uint8_t* p, uint32_t bits) {
	if (bits == 0) return 0;
	if (bits == 3) return 16;
	if (bits == 1) {
		const uint32_t w = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; // order is irrelevant for a count
		return 4 + __popc(w & (w >> 1) & 0x55555555u);
	}
	uint32_t n = 8;
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const uint32_t w = (uint32_t)p[h * 4] | (uint32_t)p[h * 4 + 1] << 8 | (uint32_t)p[h * 4 + 2] << 16 | (uint32_t)p[h * 4 + 3] << 24;
		n += __popc(w & (w >> 1) & (w >> 2) & (w >> 3) & 0x11111111u);
	}
	return n;
}

// delta i (0..15) of a group that starts at p
__device__ __forceinline__ uint32_t group_value(const uint8_t* p, uint32_t bits, uint32_t i) {
	if (bits == 0) return 0;
	if (bits == 3) return p[i];
	if (bits == 1) { // 2-bit codes, first code in the top bits of the first byte
		const uint32_t W = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | (uint32_t)p[3];
		const uint32_t c = (W >> (30 - 2 * i)) & 3u;
		if (c != 3u) return c;
		const uint32_t M = W & (W >> 1) & 0x55555555u;           // bit 30-2j set <=> code j is the sentinel
		const uint32_t before = i ? (M >> (32 - 2 * i)) : 0u;    // sentinels among codes 0..i-1
		return p[4 + __popc(before)];
	}
	unsigned long long W = 0; // 4-bit codes, high nibble first
#pragma unroll
	for (int b = 0; b < 8; ++b) W = (W << 8) | p[b];
	const uint32_t c = (uint32_t)(W >> (60 - 4 * i)) & 15u;
	if (c != 15u) return c;
	const unsigned long long M = W & (W >> 1) & (W >> 2) & (W >> 3) & 0x1111111111111111ull;
	const unsigned long long before = i ? (M >> (64 - 4 * i)) : 0ull;
	return p[8 + __popcll(before)];
}

// ---- vscan: where every plane of every block starts; meshopt_decodeVertexBuffer's framing checks and return codes ----------
// One WARP per stream.  The walk itself is a serial chain (a group's position is the sum of the sizes before it), so it is
// made as short as a chain can be: the warp stages the next kScanWin bytes of the stream in shared memory and tabulates, for
// EVERY byte offset of the window, the sentinel count a 2-bit and a 4-bit group starting there would have (32 lanes, a few
// hundred instructions); lane 0 then walks groups with one shared-memory lookup and one add per group (~40 cycles, against
// ~850 for a thread that chases bytes in global memory), records where each plane starts, and asks for the next window.
constexpr uint32_t kScanWin = 1024;                 // offsets with tabulated counts per window
constexpr uint32_t kScanWords = kScanWin / 4 + 4;   // staged words: the last offset reads 8 bytes + alignment slack
constexpr int kScanWarps = 4;
struct ScanSmem {
	uint32_t w[kScanWords];
	uint8_t c1[kScanWin], c2[kScanWin];
};
__global__ void __launch_bounds__(kScanWarps * 32) vscan_kernel(const MeshoptStream* __restrict__ streams, uint32_t nStreams, const uint8_t* __restrict__ src,
                                                               uint32_t* __restrict__ planeOff, int32_t* __restrict__ status) {
	__shared__ ScanSmem sm[kScanWarps];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t sid = blockIdx.x * kScanWarps + warp;
	if (sid >= nStreams) return;
	ScanSmem& S = sm[warp];
	const MeshoptStream st = streams[sid];
	const uint8_t* s = src + st.src_off;
	const uint32_t size = (uint32_t)st.src_size; // < 2^30 (checked when the plan is built)
	const uint32_t stride = st.stride;
	int rc = 0;
	if ((unsigned long long)size < 1ull + stride) rc = -2;
	else if ((s[0] & 0xf0) != 0xa0 || (s[0] & 0x0f) > 0) rc = -1;
	if (rc == 0) {
		// walk state (meaningful in lane 0, carried by every lane so the loop stays convergent)
		uint32_t pos = 1, v0 = 0, k = 0, g = 0xffffffffu /* at a plane header */, hbits = 0;
		uint32_t n = min(st.block_size, st.count), groups = round16(n) / 16, hs = (groups + 3) / 4;
		uint32_t* po = planeOff + st.plane_base;
		bool done = st.count == 0;
		while (!done && rc == 0) {
			// stage [base, base + 4*kScanWords) of the stream, base = pos rounded down to a 4-byte boundary of the ADDRESS
			const uintptr_t addr = (uintptr_t)(s + pos);
			const uint32_t mis = (uint32_t)(addr & 3u);
			const uint32_t* aligned = (const uint32_t*)(addr - mis);
			const uint32_t avail = size - pos + mis;     // bytes from the aligned base to the end of the stream
			for (uint32_t i = lane; i < kScanWords; i += 32) S.w[i] = (i * 4 < avail) ? __ldg(aligned + i) : 0u;
			__syncwarp();
			for (uint32_t o = lane; o < kScanWin; o += 32) {
				const uint32_t i = o >> 2, sh = (o & 3u) * 8;
				const uint32_t lo = __funnelshift_r(S.w[i], S.w[i + 1], sh), hi = __funnelshift_r(S.w[i + 1], S.w[i + 2], sh);
				S.c1[o] = (uint8_t)__popc(lo & (lo >> 1) & 0x55555555u);
				S.c2[o] = (uint8_t)(__popc(lo & (lo >> 1) & (lo >> 2) & (lo >> 3) & 0x11111111u) + __popc(hi & (hi >> 1) & (hi >> 2) & (hi >> 3) & 0x11111111u));
			}
			__syncwarp();
			if (lane == 0) {
				const uint32_t base = pos - mis; // stream offset of window byte 0
				for (;;) {
					const uint32_t o = pos - base;
					if (g == 0xffffffffu) { // a plane header: hs <= 4 bytes, 2 bits per group
						if (size - pos < hs) { rc = -2; break; }
						if (o + 4 > kScanWin) break; // next window
						*po++ = pos;
						hbits = __funnelshift_r(S.w[o >> 2], S.w[(o >> 2) + 1], (o & 3u) * 8);
						if (hs < 4) hbits &= (1u << (8 * hs)) - 1u;
						pos += hs;
						g = 0;
						continue;
					}
					if (g < groups) {
						if (size - pos < 24) { rc = -2; break; } // kByteGroupDecodeLimit
						if (o >= kScanWin) break;                 // next window
						const uint32_t bits = (hbits >> (2 * g)) & 3u;
						pos += bits == 0 ? 0u : bits == 3 ? 16u : bits == 1 ? 4u + S.c1[o] : 8u + S.c2[o];
						++g;
						continue;
					}
					// plane finished
					g = 0xffffffffu;
					if (++k == stride) {
						k = 0;
						v0 += st.block_size;
						if (v0 >= st.count) { done = true; break; }
						n = min(st.block_size, st.count - v0); groups = round16(n) / 16; hs = (groups + 3) / 4;
					}
				}
			}
			pos = __shfl_sync(0xffffffffu, pos, 0); g = __shfl_sync(0xffffffffu, g, 0); k = __shfl_sync(0xffffffffu, k, 0);
			v0 = __shfl_sync(0xffffffffu, v0, 0); hbits = __shfl_sync(0xffffffffu, hbits, 0);
			n = __shfl_sync(0xffffffffu, n, 0); groups = __shfl_sync(0xffffffffu, groups, 0); hs = __shfl_sync(0xffffffffu, hs, 0);
			rc = __shfl_sync(0xffffffffu, rc, 0); done = __shfl_sync(0xffffffffu, (int)done, 0) != 0;
			po = (uint32_t*)__shfl_sync(0xffffffffu, (unsigned long long)po, 0);
			__syncwarp();
		}
		const uint32_t tail = stride < 32 ? 32 : stride;
		if (rc == 0 && size - pos != tail) rc = -3;
	}
	if (lane == 0) status[st.view] = rc;
}

// ---- vdecode: one thread block per vertex block -----------------------------------------------------------------------------
constexpr int kVdThreads = 256;
__global__ void __launch_bounds__(kVdThreads) vdecode_kernel(const MeshoptStream* __restrict__ streams, const uint32_t* __restrict__ blockStream,
                                                            const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                            const uint32_t* __restrict__ planeOff, const int32_t* __restrict__ status,
                                                            uint32_t* __restrict__ totals) {
	__shared__ uint32_t gtab[256 * 16];   // per plane, per group: byte offset in the stream << 2 | bitslog2
	__shared__ uint32_t tile[2048];       // the decoded block, vertex-major: <= 8192 bytes (kVertexBlockSizeBytes)
	__shared__ uint32_t wsum[kVdThreads / 32];
	const uint32_t b = blockIdx.x;
	const MeshoptStream st = streams[blockStream[b]];
	if (status[st.view] != 0) return;
	const uint8_t* s = src + st.src_off;
	const uint32_t stride = st.stride, words = stride / 4;
	const uint32_t lb = b - st.first_block, v0 = lb * st.block_size;
	const uint32_t n = min(st.block_size, st.count - v0);
	const uint32_t groups = round16(n) / 16, hs = (groups + 3) / 4;
	const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;

	// every plane's groups, walked by one thread per plane (all planes side by side)
	for (uint32_t k = t; k < stride; k += kVdThreads) {
		uint32_t pos = planeOff[st.plane_base + lb * stride + k];
		const uint32_t hdr = pos;
		pos += hs;
		for (uint32_t g = 0; g < groups; ++g) {
			const uint32_t bits = (s[hdr + g / 4] >> ((g % 4) * 2)) & 3u;
			gtab[k * 16 + g] = pos << 2 | bits;
			pos += group_size(s + pos, bits);
		}
	}
	__syncthreads();

	for (uint32_t k4 = 0; k4 < words; ++k4) {
		uint32_t x = 0;
		if (t < n) {
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) {
				const uint32_t e = gtab[(k4 * 4 + j) * 16 + (t >> 4)];
				const uint32_t v = group_value(s + (e >> 2), e & 3u, t & 15u);
				x |= (unzigzag8x(v) & 0xffu) << (8 * j);
			}
		}
		// inclusive scan over the block's vertices, four byte planes per word (vertexcodec.cpp:394-404 minus the block's base)
#pragma unroll
		for (uint32_t d = 1; d < 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= d) x = __vadd4(x, y);
		}
		if (lane == 31) wsum[warp] = x;
		__syncthreads();
		uint32_t pre = 0;
		for (uint32_t w = 0; w < warp; ++w) pre = __vadd4(pre, wsum[w]);
		x = __vadd4(x, pre);
		if (t < n) tile[t * words + k4] = x;
		if (t == n - 1) totals[st.plane_base / 4 + lb * words + k4] = x;
		__syncthreads();
	}
	uint32_t* out = (uint32_t*)(dst + st.dst_off + (size_t)v0 * stride);
	for (uint32_t q = t; q < n * words; q += kVdThreads) out[q] = tile[q];
}

// ---- vcarry: per stream and word of planes, the exclusive running sum of the block totals, seeded with the tail vertex ------
__global__ void vcarry_kernel(const MeshoptStream* __restrict__ streams, uint32_t nStreams, const uint8_t* __restrict__ src,
                              const int32_t* __restrict__ status, uint32_t* __restrict__ totals /* in: totals, out: carries */) {
	const uint32_t sid = blockIdx.x;
	if (sid >= nStreams) return;
	const MeshoptStream st = streams[sid];
	if (status[st.view] != 0) return;
	const uint32_t words = st.stride / 4;
	const uint32_t nBlocks = (st.count + st.block_size - 1) / st.block_size;
	for (uint32_t k4 = threadIdx.x; k4 < words; k4 += blockDim.x) {
		const uint8_t* tail = src + st.src_off + st.src_size - st.stride + k4 * 4; // vertexcodec.cpp:1216-1217
		uint32_t base = (uint32_t)tail[0] | (uint32_t)tail[1] << 8 | (uint32_t)tail[2] << 16 | (uint32_t)tail[3] << 24;
		uint32_t* p = totals + st.plane_base / 4 + k4;
		for (uint32_t lb = 0; lb < nBlocks; ++lb, p += words) {
			const uint32_t tot = *p;
			*p = base;
			base = __vadd4(base, tot);
		}
	}
}

__global__ void __launch_bounds__(kVdThreads) vadd_kernel(const MeshoptStream* __restrict__ streams, const uint32_t* __restrict__ blockStream,
                                                         uint8_t* __restrict__ dst, const int32_t* __restrict__ status, const uint32_t* __restrict__ carries) {
	__shared__ uint32_t sc[64];
	const uint32_t b = blockIdx.x;
	const MeshoptStream st = streams[blockStream[b]];
	if (status[st.view] != 0) return;
	const uint32_t words = st.stride / 4;
	const uint32_t lb = b - st.first_block, v0 = lb * st.block_size;
	const uint32_t n = min(st.block_size, st.count - v0);
	for (uint32_t k4 = threadIdx.x; k4 < words; k4 += kVdThreads) sc[k4] = carries[st.plane_base / 4 + lb * words + k4];
	__syncthreads();
	uint32_t* out = (uint32_t*)(dst + st.dst_off + (size_t)v0 * st.stride);
	for (uint32_t q = threadIdx.x; q < n * words; q += kVdThreads) out[q] = __vadd4(out[q], sc[q % words]);
}

// ---- index codecs: one thread per stream (indexcodec.cpp) ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t vbyte(const uint8_t*& data) { // :95-120
	const uint32_t lead = *data++;
	if (lead < 128) return lead;
	uint32_t result = lead & 127u, shift = 7;
	for (int i = 0; i < 4; ++i) {
		const uint32_t g = *data++;
		result |= (g & 127u) << shift;
		shift += 7;
		if (g < 128) break;
	}
	return result;
}
__device__ __forceinline__ uint32_t delta_index(const uint8_t*& data, uint32_t last) { // :129-135
	const uint32_t v = vbyte(data);
	return last + ((v >> 1) ^ (0u - (v & 1u)));
}
__device__ __forceinline__ void put_index(uint8_t* dst, size_t i, uint32_t index_size, uint32_t v) {
	if (index_size == 2) ((unsigned short*)dst)[i] = (unsigned short)v;
	else ((uint32_t*)dst)[i] = v;
}

// the FIFOs live in local memory (L1-resident; a shared-memory layout measured 10 % slower, see profiles/README.md)
__device__ int decode_triangles(uint8_t* dst, uint32_t index_count, uint32_t index_size, const uint8_t* buffer, uint64_t buffer_size) {
	// :362-540
	if (index_count % 3 || (index_size != 2 && index_size != 4)) return -1;
	if (buffer_size < 1ull + index_count / 3 + 16) return -2;
	if ((buffer[0] & 0xf0) != 0xe0) return -1;
	const int version = buffer[0] & 0x0f;
	if (version > 1) return -1;
	uint32_t ef0[16], ef1[16], vf[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) ef0[i] = ef1[i] = vf[i] = 0xffffffffu;
	uint32_t eo = 0, vo = 0, next = 0, last = 0;
	const int fecmax = version >= 1 ? 13 : 15;
	const uint8_t* code = buffer + 1;
	const uint8_t* data = code + index_count / 3;
	const uint8_t* safe_end = buffer + buffer_size - 16;
	const uint8_t* aux = safe_end;
	for (uint32_t i = 0; i < index_count; i += 3) {
		if (data > safe_end) return -2;
		const uint32_t ct = *code++;
		uint32_t a, b, c;
		if (ct < 0xf0) {
			const uint32_t fe = ct >> 4, fec = ct & 15u;
			a = ef0[(eo - 1 - fe) & 15u];
			b = ef1[(eo - 1 - fe) & 15u];
			if ((int)fec < fecmax) {
				const uint32_t cf = vf[(vo - 1 - fec) & 15u];
				c = fec == 0 ? next : cf;
				const uint32_t fec0 = fec == 0;
				next += fec0;
				vf[vo] = c; vo = (vo + fec0) & 15u;
			} else {
				last = c = fec != 15 ? last + (fec - (fec ^ 3u)) : delta_index(data, last);
				vf[vo] = c; vo = (vo + 1) & 15u;
			}
			ef0[eo] = c; ef1[eo] = b; eo = (eo + 1) & 15u;
			ef0[eo] = a; ef1[eo] = c; eo = (eo + 1) & 15u;
		} else if (ct < 0xfe) {
			const uint32_t ca = aux[ct & 15u];
			const uint32_t feb = ca >> 4, fec = ca & 15u;
			a = next++;
			const uint32_t bf = vf[(vo - feb) & 15u];
			b = feb == 0 ? next : bf;
			const uint32_t feb0 = feb == 0;
			next += feb0;
			const uint32_t cf = vf[(vo - fec) & 15u];
			c = fec == 0 ? next : cf;
			const uint32_t fec0 = fec == 0;
			next += fec0;
			vf[vo] = a; vo = (vo + 1) & 15u;
			vf[vo] = b; vo = (vo + feb0) & 15u;
			vf[vo] = c; vo = (vo + fec0) & 15u;
			ef0[eo] = b; ef1[eo] = a; eo = (eo + 1) & 15u;
			ef0[eo] = c; ef1[eo] = b; eo = (eo + 1) & 15u;
			ef0[eo] = a; ef1[eo] = c; eo = (eo + 1) & 15u;
		} else {
			const uint32_t ca = *data++;
			const uint32_t fea = ct == 0xfe ? 0u : 15u, feb = ca >> 4, fec = ca & 15u;
			if (ca == 0) next = 0;
			a = fea == 0 ? next++ : 0u;
			b = feb == 0 ? next++ : vf[(vo - feb) & 15u];
			c = fec == 0 ? next++ : vf[(vo - fec) & 15u];
			if (fea == 15) last = a = delta_index(data, last);
			if (feb == 15) last = b = delta_index(data, last);
			if (fec == 15) last = c = delta_index(data, last);
			vf[vo] = a; vo = (vo + 1) & 15u;
			vf[vo] = b; vo = (vo + ((feb == 0) | (feb == 15))) & 15u;
			vf[vo] = c; vo = (vo + ((fec == 0) | (fec == 15))) & 15u;
			ef0[eo] = b; ef1[eo] = a; eo = (eo + 1) & 15u;
			ef0[eo] = c; ef1[eo] = b; eo = (eo + 1) & 15u;
			ef0[eo] = a; ef1[eo] = c; eo = (eo + 1) & 15u;
		}
		put_index(dst, i, index_size, a); put_index(dst, i + 1, index_size, b); put_index(dst, i + 2, index_size, c);
	}
	return data == safe_end ? 0 : -3;
}

__device__ int decode_sequence(uint8_t* dst, uint32_t index_count, uint32_t index_size, const uint8_t* buffer, uint64_t buffer_size) {
	// :618-672
	if (index_size != 2 && index_size != 4) return -1;
	if (buffer_size < 1ull + index_count + 4) return -2;
	if ((buffer[0] & 0xf0) != 0xd0) return -1;
	if ((buffer[0] & 0x0f) > 1) return -1;
	const uint8_t* data = buffer + 1;
	const uint8_t* safe_end = buffer + buffer_size - 4;
	uint32_t last0 = 0, last1 = 0;
	for (uint32_t i = 0; i < index_count; ++i) {
		if (data >= safe_end) return -2;
		uint32_t v = vbyte(data);
		const uint32_t cur = v & 1u;
		v >>= 1;
		const uint32_t index = (cur ? last1 : last0) + ((v >> 1) ^ (0u - (v & 1u)));
		if (cur) last1 = index; else last0 = index;
		put_index(dst, i, index_size, index);
	}
	return data == safe_end ? 0 : -3;
}

__global__ void idecode_kernel(const MeshoptStream* __restrict__ streams, uint32_t nStreams, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                               int32_t* __restrict__ status) {
	const uint32_t sid = blockIdx.x * blockDim.x + threadIdx.x;
	if (sid >= nStreams) return;
	const MeshoptStream st = streams[sid];
	status[st.view] = st.mode == 1 ? decode_triangles(dst + st.dst_off, st.count, st.stride, src + st.src_off, st.src_size)
	                               : decode_sequence(dst + st.dst_off, st.count, st.stride, src + st.src_off, st.src_size);
}

// ---- filters, scalar definitions (vertexfilter.cpp:75-160); compiled with -fmad=false, IEEE sqrt and division ------------------
__device__ __forceinline__ int round_half_away(float v, float ref) { return (int)(v + (ref >= 0.f ? 0.5f : -0.5f)); }

__global__ void filter_kernel(const MeshoptStream* __restrict__ filtered, uint32_t nFiltered, const unsigned long long* __restrict__ elemFirst,
                              unsigned long long total, uint8_t* __restrict__ dst, const int32_t* __restrict__ status) {
	for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (unsigned long long)gridDim.x * blockDim.x) {
		uint32_t lo = 0, hi = nFiltered - 1; // last view with elemFirst <= e
		while (lo < hi) {
			const uint32_t mid = (lo + hi + 1) >> 1;
			if (elemFirst[mid] <= e) lo = mid; else hi = mid - 1;
		}
		const MeshoptStream st = filtered[lo];
		if (status[st.view] != 0) continue;
		const unsigned long long i = e - elemFirst[lo];
		uint8_t* base = dst + st.dst_off;
		if (st.filter == 3) { // exponential: one 32-bit word per element
			uint32_t* p = (uint32_t*)base + i;
			const uint32_t v = *p;
			const int m = (int)(v << 8) >> 8, ex = (int)v >> 24;
			*p = __float_as_uint(__uint_as_float((uint32_t)(ex + 127) << 23) * (float)m);
		} else if (st.filter == 1) { // octahedral, 4 x int8 or 4 x int16 per element
			float x, y, zc, mx;
			if (st.stride == 4) { const signed char* d = (const signed char*)base + i * 4; x = d[0]; y = d[1]; zc = d[2]; mx = 127.f; }
			else { const short* d = (const short*)base + i * 4; x = d[0]; y = d[1]; zc = d[2]; mx = 32767.f; }
			const float z = zc - fabsf(x) - fabsf(y);
			const float t = (z >= 0.f) ? 0.f : z;
			x += (x >= 0.f) ? t : -t;
			y += (y >= 0.f) ? t : -t;
			const float l = sqrtf(x * x + y * y + z * z);
			const float s = mx / l;
			const int xf = round_half_away(x * s, x), yf = round_half_away(y * s, y), zf = round_half_away(z * s, z);
			if (st.stride == 4) { signed char* d = (signed char*)base + i * 4; d[0] = (signed char)xf; d[1] = (signed char)yf; d[2] = (signed char)zf; }
			else { short* d = (short*)base + i * 4; d[0] = (short)xf; d[1] = (short)yf; d[2] = (short)zf; }
		} else if (st.filter == 2) { // quaternion, 4 x int16
			short* d = (short*)base + i * 4;
			const float scale = 1.f / sqrtf(2.f);
			const int c3 = d[3];
			const float ss = scale / (float)(c3 | 3);
			const float x = (float)d[0] * ss, y = (float)d[1] * ss, z = (float)d[2] * ss;
			const float ww = 1.f - x * x - y * y - z * z;
			const float w = sqrtf(ww >= 0.f ? ww : 0.f);
			const int xf = round_half_away(x * 32767.f, x), yf = round_half_away(y * 32767.f, y), zf = round_half_away(z * 32767.f, z);
			const int wf = (int)(w * 32767.f + 0.5f);
			const int qc = c3 & 3;
			d[(qc + 1) & 3] = (short)xf; d[(qc + 2) & 3] = (short)yf; d[(qc + 3) & 3] = (short)zf; d[qc & 3] = (short)wf;
		}
	}
}

} // namespace

cudaError_t launch_meshopt_decode(const MeshoptPlan& p, const uint8_t* src, uint8_t* dst, int num_sms, cudaStream_t stream, int* launches) {
	// (running the index kernel on a second stream beside the vertex kernels was measured: the single-thread chains get starved
	// by the wide vertex blocks and the whole decode takes 2.8x longer — one stream, back to back)
	int n = 0;
	if (p.nIndexStreams) { idecode_kernel<<<(p.nIndexStreams + 31) / 32, 32, 0, stream>>>(p.indexStreams, p.nIndexStreams, src, dst, p.status); ++n; }
	if (p.nVertexStreams) {
		vscan_kernel<<<(p.nVertexStreams + kScanWarps - 1) / kScanWarps, kScanWarps * 32, 0, stream>>>(p.vertexStreams, p.nVertexStreams, src, p.planeOff, p.status); ++n;
		if (p.nBlocks) {
			vdecode_kernel<<<p.nBlocks, kVdThreads, 0, stream>>>(p.vertexStreams, p.blockStream, src, dst, p.planeOff, p.status, p.totals); ++n;
			vcarry_kernel<<<p.nVertexStreams, 64, 0, stream>>>(p.vertexStreams, p.nVertexStreams, src, p.status, p.totals); ++n;
			vadd_kernel<<<p.nBlocks, kVdThreads, 0, stream>>>(p.vertexStreams, p.blockStream, dst, p.status, p.totals); ++n;
		}
	}
	if (p.nFiltered && p.filterElems) {
		unsigned long long grid = (p.filterElems + 255) / 256;
		if (grid > (unsigned long long)num_sms * 16) grid = (unsigned long long)num_sms * 16;
		filter_kernel<<<(unsigned)grid, 256, 0, stream>>>(p.filtered, p.nFiltered, p.elemFirst, p.filterElems, dst, p.status); ++n;
	}
	if (launches) *launches = n;
	return cudaGetLastError();
}
