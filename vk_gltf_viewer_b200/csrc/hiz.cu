// hiz.cu — depth-pyramid build from the 64-bit visbuffer.
// Replaces shaders/hiz_reduce.comp.glsl:21-31 and its per-mip dispatch loop (application.cpp:951-1003): the reference
// issues one 32x32-group dispatch per mip with a full barrier between mips (9-12 serialised launches); here the
// leading "exact 2x" mips are produced by ONE tiled launch straight from the visbuffer (depth extraction fused into
// the level-0 read), and the remaining small mips by ONE single-block launch that restates the min-sampler footprint
// rule texel by texel (i0 = floor(u*S - 0.5), {i0, i0+1} minus zero-weight texels, CLAMP_TO_EDGE).
#include "kernels.cuh"

namespace {

constexpr int kTileW = 64, kTileH = 16; // source pixels per warp tile; yields 32x8, 16x4, 8x2, 4x1 texels of mips 0..3
constexpr int kHizWarps = 8;

__device__ __forceinline__ float min4(float a, float b, float c, float d) { return gmin(gmin(gmin(a, b), c), d); }

// One warp per 64x16 source tile. Lane l owns source columns 2l,2l+1 (one 16-byte load per row, 16 loads in flight).
// Valid only for levels whose source is exactly twice the destination in both axes: the sampler footprint is then the
// aligned 2x2 quad {2p, 2p+1} (u = 2p + 0.5 up to rounding noise << 0.5; checked exhaustively in tests/test_hiz_rule.py).
__global__ void __launch_bounds__(kHizWarps * 32) hiz_tiled_kernel(const HizParams p) {
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t tilesX = (p.W + kTileW - 1) / kTileW, tilesY = (p.H + kTileH - 1) / kTileH;
	const uint32_t nTiles = tilesX * tilesY;
	const uint32_t E = p.exact_levels;
	for (uint32_t tile = blockIdx.x * kHizWarps + (threadIdx.x >> 5); tile < nTiles; tile += gridDim.x * kHizWarps) {
		const uint32_t tx = tile % tilesX, ty = tile / tilesX;
		const uint32_t x0 = tx * kTileW + lane * 2, y0 = ty * kTileH;
		const bool colIn = x0 < p.W; // W is even whenever E >= 1, so the pair is in or out together
		float m0[8];
		{
			ulonglong2 v[kTileH];
#pragma unroll
			for (int r = 0; r < kTileH; ++r) {
				v[r] = make_ulonglong2(0ull, 0ull); // key 0 == depth +NaN pattern never produced; replaced below by +inf
				if (colIn && y0 + r < p.H) v[r] = __ldcs((const ulonglong2*)(p.vis + (size_t)(y0 + r) * p.W + x0));
			}
#pragma unroll
			for (int r = 0; r < 8; ++r) {
				const bool in = colIn && (y0 + 2 * r + 1 < p.H);
				const float a = depth_of_key(v[2 * r].x), b = depth_of_key(v[2 * r].y);
				const float c = depth_of_key(v[2 * r + 1].x), d = depth_of_key(v[2 * r + 1].y);
				m0[r] = in ? min4(a, b, c, d) : __int_as_float(0x7f800000);
			}
		}
		// mip 0: 32 x 8 per tile
		{
			float* dst = p.pyramid + p.pyr.off[0];
			const uint32_t mw = p.pyr.w[0], mx = tx * 32 + lane, my0 = ty * 8;
#pragma unroll
			for (int r = 0; r < 8; ++r)
				if (mx < (p.W >> 1) && my0 + r < (p.H >> 1)) dst[(size_t)(my0 + r) * mw + mx] = m0[r];
		}
		if (E < 2) continue;
		float m1[4];
#pragma unroll
		for (int r = 0; r < 4; ++r) {
			const float v = gmin(m0[2 * r], m0[2 * r + 1]);
			m1[r] = gmin(v, __shfl_xor_sync(0xffffffffu, v, 1));
		}
		if ((lane & 1) == 0) {
			float* dst = p.pyramid + p.pyr.off[1];
			const uint32_t mw = p.pyr.w[1], mx = tx * 16 + (lane >> 1), my0 = ty * 4;
#pragma unroll
			for (int r = 0; r < 4; ++r)
				if (mx < (p.W >> 2) && my0 + r < (p.H >> 2)) dst[(size_t)(my0 + r) * mw + mx] = m1[r];
		}
		if (E < 3) continue;
		float m2[2];
#pragma unroll
		for (int r = 0; r < 2; ++r) {
			const float v = gmin(m1[2 * r], m1[2 * r + 1]);
			m2[r] = gmin(v, __shfl_xor_sync(0xffffffffu, v, 2));
		}
		if ((lane & 3) == 0) {
			float* dst = p.pyramid + p.pyr.off[2];
			const uint32_t mw = p.pyr.w[2], mx = tx * 8 + (lane >> 2), my0 = ty * 2;
#pragma unroll
			for (int r = 0; r < 2; ++r)
				if (mx < (p.W >> 3) && my0 + r < (p.H >> 3)) dst[(size_t)(my0 + r) * mw + mx] = m2[r];
		}
		if (E < 4) continue;
		{
			const float v = gmin(m2[0], m2[1]);
			const float m3 = gmin(v, __shfl_xor_sync(0xffffffffu, v, 4));
			if ((lane & 7) == 0) {
				float* dst = p.pyramid + p.pyr.off[3];
				const uint32_t mw = p.pyr.w[3], mx = tx * 4 + (lane >> 3), my = ty;
				if (mx < (p.W >> 4) && my < (p.H >> 4)) dst[(size_t)my * mw + mx] = m3;
			}
		}
	}
}

// Remaining mips, one block, level after level (each level is a few thousand texels at most and depends on the previous).
// first_level = index of the first pyramid mip to produce here (>= 1 unless the depth image is tiny).
__global__ void __launch_bounds__(1024) hiz_tail_kernel(const HizParams p, uint32_t first_level) {
	for (uint32_t k = first_level; k < p.pyr.levels; ++k) {
		const uint32_t i = k + 1;                       // reference view index of the destination (application.cpp:964)
		const uint32_t dw = p.W >> i, dh = p.H >> i;    // levelSize
		if (dw == 0 || dh == 0) continue;               // zero-sized dispatch: mip keeps its contents
		float* dst = p.pyramid + p.pyr.off[k];
		const uint32_t dstride = p.pyr.w[k];
		if (k == 0) {
			// source = depth image itself (only when no level was exact, e.g. odd W/H): read depth from the visbuffer keys
			for (uint32_t t = threadIdx.x; t < dw * dh; t += blockDim.x) {
				const uint32_t x = t % dw, y = t / dw;
				const float u = ((float)x + 0.5f) / (float)dw, v = ((float)y + 0.5f) / (float)dh;
				int x0, x1, y0, y1;
				footprint(u, p.W, x0, x1);
				footprint(v, p.H, y0, y1);
				float m = depth_of_key(p.vis[(size_t)y0 * p.W + x0]);
				m = gmin(m, depth_of_key(p.vis[(size_t)y0 * p.W + x1]));
				m = gmin(m, depth_of_key(p.vis[(size_t)y1 * p.W + x0]));
				m = gmin(m, depth_of_key(p.vis[(size_t)y1 * p.W + x1]));
				dst[(size_t)y * dstride + x] = m;
			}
		} else {
			const float* src = p.pyramid + p.pyr.off[k - 1];
			const uint32_t sw = p.pyr.w[k - 1], sh = p.pyr.h[k - 1];
			for (uint32_t t = threadIdx.x; t < dw * dh; t += blockDim.x) {
				const uint32_t x = t % dw, y = t / dw;
				// hiz_reduce.comp.glsl:28 : texture(src, (vec2(pos) + 0.5) / imageSize)
				const float u = ((float)x + 0.5f) / (float)dw, v = ((float)y + 0.5f) / (float)dh;
				int x0, x1, y0, y1;
				footprint(u, sw, x0, x1);
				footprint(v, sh, y0, y1);
				// plain (coherent) loads: the source was written by this block in the previous iteration
				float m = src[(size_t)y0 * sw + x0];
				m = gmin(m, src[(size_t)y0 * sw + x1]);
				m = gmin(m, src[(size_t)y1 * sw + x0]);
				m = gmin(m, src[(size_t)y1 * sw + x1]);
				dst[(size_t)y * dstride + x] = m;
			}
		}
		__syncthreads();
	}
}

} // namespace

cudaError_t launch_hiz(const HizParams& p, int num_sms, cudaStream_t stream, int* launches) {
	uint32_t first_tail = 0;
	if (p.exact_levels >= 1) {
		const uint32_t tiles = ((p.W + kTileW - 1) / kTileW) * ((p.H + kTileH - 1) / kTileH);
		uint32_t grid = (tiles + kHizWarps - 1) / kHizWarps;
		if (grid > (uint32_t)num_sms * 8) grid = (uint32_t)num_sms * 8;
		hiz_tiled_kernel<<<grid, kHizWarps * 32, 0, stream>>>(p);
		if (launches) ++*launches;
		first_tail = p.exact_levels;
	}
	if (first_tail < p.pyr.levels) {
		hiz_tail_kernel<<<1, 1024, 0, stream>>>(p, first_tail);
		if (launches) ++*launches;
	}
	return cudaGetLastError();
}
