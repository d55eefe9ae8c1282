// hiz.cu — depth-pyramid build from the 64-bit visbuffer.
// Replaces shaders/hiz_reduce.comp.glsl:21-31 and its per-mip dispatch loop (application.cpp:951-1003): the reference
// issues one 32x32-group dispatch per mip with a full barrier between mips (9-12 serialised launches); here the
// leading "exact 2x" mips are produced by ONE tiled launch straight from the visbuffer (depth extraction fused into
// the level-0 read); the block of that launch that finishes LAST then produces the remaining small mips, staged in shared
// memory level after level, restating the min-sampler footprint rule texel by texel (i0 = floor(u*S - 0.5),
// {i0, i0+1} minus zero-weight texels, CLAMP_TO_EDGE).  One launch per pyramid build.
#include "kernels.cuh"
#include "hiz_tile.cuh"
#include "xgpu.cuh"
#ifdef VKV_HIZ_DEBUG
#include <cstdio>
#define TS(i) do { if (threadIdx.x == 0) dbg_ts[i] = clock64(); } while (0)
#else
#define TS(i)
#endif


namespace {

constexpr int kHizWarps = 32; // 1024 threads: the last block also runs the serial small-mip tail, which wants the whole SM

// Remaining mips, ONE block, level after level (each level is a few thousand texels at most and depends on the previous).
// first_level = index of the first pyramid mip to produce here.  Everything the serial chain touches lives in shared memory:
//   * the pyramid geometry (dynamic indexing of kernel parameters costs a constant-cache miss per level),
//   * the separable sampler footprints of ALL tail levels, evaluated up front (columns depend on x only, rows on y only),
//   * the source mip of the first tail level, staged with coalesced 16-byte L2 loads when it fits (4K / 1080p: 240x135),
//   * every produced level (two ping-pong buffers; levels shrink 4x).
constexpr uint32_t kTailSrc = 33792;                     // floats: staged source of the first tail level (240x135 = 32400)
constexpr uint32_t kTailBufA = 8192, kTailBufB = 2304;   // floats: 120x67 and 60x33
constexpr uint32_t kFpCap = 1024;                        // ushort2 entries: column + row footprints of all tail levels
constexpr size_t kTailSmemBytes = (kTailSrc + kTailBufA + kTailBufB) * 4 + kFpCap * 4 + 64 * 4;

struct TailSmem {
	float* src; float* bufA; float* bufB; ushort2* fp; uint32_t* geo;
	__device__ explicit TailSmem(unsigned char* base) {
		src = (float*)base; bufA = src + kTailSrc; bufB = bufA + kTailBufA; fp = (ushort2*)(bufB + kTailBufB); geo = (uint32_t*)(fp + kFpCap);
	}
};

__device__ __forceinline__ void hiz_tail(const HizParams& p, uint32_t first_level, const TailSmem& sm) {
#ifdef VKV_HIZ_DEBUG
	__shared__ long long dbg_ts[24];
#endif
	TS(0);
	// geometry -> shared: geo[k] = off, geo[16+k] = w, geo[32+k] = h, geo[48+k] = footprint table offset of level k
	if (threadIdx.x < 16) {
		sm.geo[threadIdx.x] = p.pyr.off[threadIdx.x];
		sm.geo[16 + threadIdx.x] = p.pyr.w[threadIdx.x];
		sm.geo[32 + threadIdx.x] = p.pyr.h[threadIdx.x];
	}
	if (threadIdx.x == 0) { // table offsets; 0xffffffff = level does not fit (evaluated per texel instead)
		uint32_t o = 0;
		for (uint32_t k = 0; k < 16; ++k) {
			const uint32_t dw = p.W >> (k + 1), dh = p.H >> (k + 1);
			if (k >= first_level && k >= 1 && k < p.pyr.levels && dw && dh && o + dw + dh <= kFpCap) { sm.geo[48 + k] = o; o += dw + dh; }
			else sm.geo[48 + k] = 0xffffffffu;
		}
	}
	__syncthreads();
	TS(1);
	const uint32_t levels = p.pyr.levels;
	{ // footprint tables of all tail levels: 4 warps per level, so the per-level latency chains run side by side
		const uint32_t wid = threadIdx.x >> 5, part = wid & 3u, groups = max(1u, blockDim.x >> 7);
		for (uint32_t k = (first_level < 1 ? 1 : first_level) + (wid >> 2); k < levels; k += groups) {
			const uint32_t fo = sm.geo[48 + k];
			if (fo == 0xffffffffu) continue;
			const uint32_t dw = p.W >> (k + 1), dh = p.H >> (k + 1), sw = sm.geo[16 + k - 1], sh = sm.geo[32 + k - 1];
			// hiz_reduce.comp.glsl:28: uv = (pos + 0.5) / imageSize, through the min-sampler footprint rule
			for (uint32_t e = part * 32 + (threadIdx.x & 31); e < dw + dh; e += 128) {
				const bool col = e < dw;
				const uint32_t pos = col ? e : e - dw;
				int lo, hi;
				footprint(((float)pos + 0.5f) / (float)(col ? dw : dh), col ? sw : sh, lo, hi);
				sm.fp[fo + e] = make_ushort2((unsigned short)lo, (unsigned short)hi);
			}
		}
	}
	TS(2);
	// stage the first level's source
	const float* ssrc = nullptr;   // shared-memory copy of mip k-1 (row stride = its width), or NULL
	if (first_level >= 1 && first_level < levels) {
		const uint32_t sw = sm.geo[16 + first_level - 1], sh = sm.geo[32 + first_level - 1], n = sw * sh, off = sm.geo[first_level - 1];
		if (n <= kTailSrc && (off & 3) == 0) {
			const float4* g = (const float4*)(p.pyramid + off);
			for (uint32_t t = threadIdx.x; t < (n >> 2); t += blockDim.x) ((float4*)sm.src)[t] = __ldcg(g + t); // written by other blocks: L2
			for (uint32_t t = (n & ~3u) + threadIdx.x; t < n; t += blockDim.x) sm.src[t] = __ldcg(p.pyramid + off + t);
			ssrc = sm.src;
		}
	}
	__syncthreads();
	TS(3);

	float* cur = sm.bufA; uint32_t curCap = kTailBufA;
	float* oth = sm.bufB; uint32_t othCap = kTailBufB;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
	for (uint32_t k = first_level; k < levels; ++k) {
		const uint32_t i = k + 1;                       // reference view index of the destination (application.cpp:964)
		const uint32_t dw = p.W >> i, dh = p.H >> i;    // levelSize
		if (dw == 0 || dh == 0) continue;               // zero-sized dispatch: mip keeps its contents
		float* dst = p.pyramid + sm.geo[k];
		const uint32_t dstride = sm.geo[16 + k];
		float* sdst = (dw * dh <= curCap && dstride == dw) ? cur : nullptr;
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
				if (sdst) sdst[t] = m;
			}
		} else {
			const float* gsrc = p.pyramid + sm.geo[k - 1];
			const uint32_t sw = sm.geo[16 + k - 1], sh = sm.geo[32 + k - 1];
			const uint32_t fo = sm.geo[48 + k];
			if (fo != 0xffffffffu) {
				const ushort2* fpCols = sm.fp + fo;
				const ushort2* fpRows = sm.fp + fo + dw;
				for (uint32_t y = warp; y < dh; y += nwarps) {
					const ushort2 ry = fpRows[y];
					const uint32_t r0 = ry.x * sw, r1 = ry.y * sw;
					// up to 4 column groups per pass: all gathers first, then all stores (the shared-memory stores would otherwise
					// serialise against the next group's loads)
					for (uint32_t xb = 0; xb < dw; xb += 128) {
						float m[4];
#pragma unroll
						for (int j = 0; j < 4; ++j) {
							const uint32_t x = xb + j * 32 + lane;
							m[j] = 0.f;
							if (x < dw) {
								const ushort2 cx = fpCols[x];
								float a, b, c, d;
								if (ssrc) { a = ssrc[r0 + cx.x]; b = ssrc[r0 + cx.y]; c = ssrc[r1 + cx.x]; d = ssrc[r1 + cx.y]; }
								else { // L2 loads (source too large to stage, or written by this block one level ago)
									a = __ldcg(gsrc + r0 + cx.x); b = __ldcg(gsrc + r0 + cx.y); c = __ldcg(gsrc + r1 + cx.x); d = __ldcg(gsrc + r1 + cx.y);
								}
								m[j] = gmin(gmin(gmin(a, b), c), d);
							}
						}
#pragma unroll
						for (int j = 0; j < 4; ++j) {
							const uint32_t x = xb + j * 32 + lane;
							if (x < dw) {
								dst[(size_t)y * dstride + x] = m[j];
								if (sdst) sdst[y * dw + x] = m[j];
							}
						}
					}
				}
			} else { // very large tail level (only for odd resolutions): per-texel evaluation
				for (uint32_t t = threadIdx.x; t < dw * dh; t += blockDim.x) {
					const uint32_t x = t % dw, y = t / dw;
					const float u = ((float)x + 0.5f) / (float)dw, v = ((float)y + 0.5f) / (float)dh;
					int x0, x1, y0, y1;
					footprint(u, sw, x0, x1);
					footprint(v, sh, y0, y1);
					float a, b, c, d;
					if (ssrc) { a = ssrc[y0 * sw + x0]; b = ssrc[y0 * sw + x1]; c = ssrc[y1 * sw + x0]; d = ssrc[y1 * sw + x1]; }
					else {
						a = __ldcg(gsrc + (size_t)y0 * sw + x0); b = __ldcg(gsrc + (size_t)y0 * sw + x1);
						c = __ldcg(gsrc + (size_t)y1 * sw + x0); d = __ldcg(gsrc + (size_t)y1 * sw + x1);
					}
					const float m = gmin(gmin(gmin(a, b), c), d);
					dst[(size_t)y * dstride + x] = m;
					if (sdst) sdst[t] = m;
				}
			}
		}
		__syncthreads();
		TS(4 + k);
		ssrc = sdst;
		if (sdst) { float* tp = cur; cur = oth; oth = tp; const uint32_t tc = curCap; curCap = othCap; othCap = tc; }
	}
#ifdef VKV_HIZ_DEBUG
	if (threadIdx.x == 0) {
		printf("geo %lld tables %lld stage %lld |", dbg_ts[1] - dbg_ts[0], dbg_ts[2] - dbg_ts[1], dbg_ts[3] - dbg_ts[2]);
		long long prev = dbg_ts[3];
		for (uint32_t k = first_level; k < levels; ++k) { printf(" L%u %lld", k, dbg_ts[4 + k] - prev); prev = dbg_ts[4 + k]; }
		printf("\n");
	}
#endif
}

// tail-only launch: resolutions without any exact level (odd W or H)
// strip mode: the exact mips this tail reads were stored by ALL ranks' strip kernels — wait for their signals first
__device__ __forceinline__ void tail_wait_for_peers(const HizParams& p) {
	if (!p.wait_flags) return;
	if ((int)threadIdx.x < p.wait_ranks) xgpu_wait(p.wait_flags, threadIdx.x, p.wait_epoch, p.wait_timeout_ns, p.wait_error);
	__syncthreads();
}

__global__ void __launch_bounds__(1024) hiz_tail_kernel(const HizParams p, uint32_t first_level) {
	extern __shared__ __align__(16) unsigned char tailSmem[];
	tail_wait_for_peers(p);
	hiz_tail(p, first_level, TailSmem(tailSmem));
}

// Tail whose FIRST level is too large for one block to take in its stride: at 8K the first non-exact mip is 240x135 texels read
// from a 480x270 source that does not fit the shared-memory stage — 32 texels per thread, each four dependent L2 loads, ~25 us for
// one block.  Here that level is spread over the grid (the generic sampler rule per texel, the arithmetic of hiz_tail's per-texel
// branch); the block that finishes last runs the remaining small levels as before.
__global__ void __launch_bounds__(1024) hiz_tail_spread_kernel(const HizParams p, uint32_t first_level) {
	extern __shared__ __align__(16) unsigned char tailSmem[];
	__shared__ uint32_t sLast;
	tail_wait_for_peers(p);
	const uint32_t k = first_level, dw = p.W >> (k + 1), dh = p.H >> (k + 1);
	if (dw && dh) {
		const float* gsrc = p.pyramid + p.pyr.off[k - 1];
		const uint32_t sw = p.pyr.w[k - 1], sh = p.pyr.h[k - 1], dstride = p.pyr.w[k];
		float* dst = p.pyramid + p.pyr.off[k];
		for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < dw * dh; t += gridDim.x * blockDim.x) {
			const uint32_t x = t % dw, y = t / dw;
			const float u = ((float)x + 0.5f) / (float)dw, v = ((float)y + 0.5f) / (float)dh; // hiz_reduce.comp.glsl:28
			int x0, x1, y0, y1;
			footprint(u, sw, x0, x1);
			footprint(v, sh, y0, y1);
			const float a = __ldcg(gsrc + (size_t)y0 * sw + x0), b = __ldcg(gsrc + (size_t)y0 * sw + x1);
			const float c = __ldcg(gsrc + (size_t)y1 * sw + x0), d = __ldcg(gsrc + (size_t)y1 * sw + x1);
			dst[(size_t)y * dstride + x] = gmin(gmin(gmin(a, b), c), d);
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		const uint32_t ticket = atomicAdd(p.done, 1u);
		sLast = (ticket == gridDim.x - 1) ? 1u : 0u;
		if (sLast) { *p.done = 0u; __threadfence(); }
	}
	__syncthreads();
	if (!sLast) return;
	hiz_tail(p, first_level + 1, TailSmem(tailSmem));
}

// One warp per 64x16 source tile (hiz_tile.cuh).  Valid only for levels whose source is exactly twice the destination in both axes:
// the sampler footprint is then the aligned 2x2 quad {2p, 2p+1} (u = 2p + 0.5 up to rounding noise << 0.5; checked exhaustively
// in tests/test_oracle.py).
__global__ void __launch_bounds__(kHizWarps * 32) hiz_tiled_kernel(const HizParams p) {
	extern __shared__ __align__(16) unsigned char tailSmem[]; // used by the last block only (1 block / SM anyway: 1024 threads)
	__shared__ uint32_t sLast;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t tilesX = (p.W + kTileW - 1) / kTileW, tilesY = (p.H + kTileH - 1) / kTileH;
	const uint32_t nTiles = tilesX * tilesY;
	// launched with programmatic stream serialization behind the raster's drain kernel (vkv_frame): set up early, start when it is done
	asm volatile("griddepcontrol.wait;" ::: "memory");
	const HizTileGeo geo = {p.W, p.H, p.exact_levels, {p.pyr.off[0], p.pyr.off[1], p.pyr.off[2], p.pyr.off[3]}, {p.pyr.w[0], p.pyr.w[1], p.pyr.w[2], p.pyr.w[3]}};
	float* const pyramid = p.pyramid;
	const uint32_t firstTile = blockIdx.x * kHizWarps + (threadIdx.x >> 5), tileStride = gridDim.x * kHizWarps;
	// Pass-B rebuild (vkv_frame, two-pass): only tiles the pass-B rasteriser marked can differ from what the pass-A build stored a few
	// launches ago.  Lane k fetches the flag of the warp's k-th tile (one round trip for all of them), the ballot says which to redo.
	uint32_t redo = 0xffffffffu;
	if (p.tile_dirty && firstTile + 31ull * tileStride >= nTiles && __ldcg(p.dirty_count) <= p.dirty_limit) { // (a warp never has more than 32 tiles below 16K x 16K; beyond, rebuild all)
		const uint32_t t = firstTile + lane * tileStride;
		redo = __ballot_sync(0xffffffffu, t < nTiles && __ldcg(p.tile_dirty + t) != 0);
	}
	if (p.tiles_done) { // statistics for the byte accounting of bench.py: one reduction per warp that has work (few in a partial rebuild)
		uint32_t mine = 0;
		if (redo != 0xffffffffu) mine = __popc(redo);
		else if (firstTile < nTiles) mine = (nTiles - firstTile + tileStride - 1) / tileStride;
		if (lane == 0 && mine) atomicAdd(p.tiles_done, mine);
	}
	uint32_t k = 0;
	for (uint32_t tile = firstTile; tile < nTiles; tile += tileStride, ++k) {
		if (!((redo >> (k & 31u)) & 1u)) continue;
		const uint32_t tx = tile % tilesX, ty = tile / tilesX;
		ulonglong2 v[kTileH];
		hiz_tile_load(p.vis, geo, tx, ty, lane, v);
		hiz_tile_reduce(v, geo, tx, ty, lane, [&](int, uint32_t idx, float m) { pyramid[idx] = m; });
	}
	// Programmatic dependent launch: this block's tiles are written; once every block has said so (or exited) the next kernel in the
	// stream may START if it was launched with programmatic stream serialization (vkv_frame: the pass-B cull, which does not touch
	// the pyramid before its griddepcontrol.wait) — the serial small-mip tail below then overlaps that kernel's front half.
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	// the block that finishes last produces the small mips (no second launch): classic last-block-done hand-off
	if (p.exact_levels >= p.pyr.levels || !p.done) return;
	// bar.sync orders every thread's mip stores before thread 0's gpu-scope fence (fences are cumulative), so ONE fence per
	// block publishes the block's tiles; a fence in all 1024 threads costs several microseconds of L1 invalidations
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		const uint32_t ticket = atomicAdd(p.done, 1u);
		sLast = (ticket == gridDim.x - 1) ? 1u : 0u;
		if (sLast) { *p.done = 0u; __threadfence(); } // ready for the next launch; acquire side of the hand-off
	}
	__syncthreads();
	if (!sLast) return;
	hiz_tail(p, p.exact_levels, TailSmem(tailSmem));
}

} // namespace

// does the first tail level overwhelm a single block?  (its source does not fit the shared-memory stage: 8K and up)
static bool tail_wants_spreading(const HizParams& p) {
	const uint32_t k = p.exact_levels;
	return k >= 1 && k < p.pyr.levels && p.done != nullptr && (size_t)p.pyr.w[k - 1] * p.pyr.h[k - 1] > kTailSrc;
}
static void launch_tail(const HizParams& p, cudaStream_t stream) {
	if (tail_wants_spreading(p)) hiz_tail_spread_kernel<<<32, 1024, kTailSmemBytes, stream>>>(p, p.exact_levels);
	else hiz_tail_kernel<<<1, 1024, kTailSmemBytes, stream>>>(p, p.exact_levels);
}

static cudaError_t launch_tiled(const HizParams& p, uint32_t grid, cudaStream_t stream) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kHizWarps * 32); cfg.dynamicSmemBytes = kTailSmemBytes; cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, hiz_tiled_kernel, p);
}

cudaError_t launch_hiz(const HizParams& p, int num_sms, cudaStream_t stream, int* launches) {
	// function attributes are per device: a process may hold contexts on several GPUs (vkv_create(device = k))
	static bool attrSet[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	bool& attr = attrSet[dev & 63];
	if (!attr) {
		cudaFuncSetAttribute(hiz_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemBytes);
		cudaFuncSetAttribute(hiz_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemBytes);
		cudaFuncSetAttribute(hiz_tail_spread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemBytes);
		attr = true;
	}
	if (p.exact_levels >= 1) {
		const uint32_t tiles = ((p.W + kTileW - 1) / kTileW) * ((p.H + kTileH - 1) / kTileH);
		uint32_t grid = (tiles + kHizWarps - 1) / kHizWarps;
		if (grid > (uint32_t)num_sms) grid = (uint32_t)num_sms;
		if (tail_wants_spreading(p)) { // 8K and up: tiles without the in-kernel tail, then the tail with its first level spread over 32 blocks
			HizParams q = p;
			q.done = nullptr;
			launch_tiled(q, grid, stream);
			launch_tail(p, stream);
			if (launches) *launches += 2;
			return cudaGetLastError();
		}
		launch_tiled(p, grid, stream); // its last block runs the tail
		if (launches) ++*launches;
		if (!p.done && p.split_tail && p.exact_levels < p.pyr.levels) { // diagnosis: tail as a second launch
			hiz_tail_kernel<<<1, 1024, kTailSmemBytes, stream>>>(p, p.exact_levels);
			if (launches) ++*launches;
		}
	} else if (p.pyr.levels) {
		hiz_tail_kernel<<<1, 1024, kTailSmemBytes, stream>>>(p, 0);
		if (launches) ++*launches;
	}
	return cudaGetLastError();
}

// the small mips alone (strip mode: every rank rebuilds them from the all-gathered last exact mip)
cudaError_t launch_hiz_tail(const HizParams& p, cudaStream_t stream) {
	static bool attrSet[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (!attrSet[dev & 63]) {
		cudaFuncSetAttribute(hiz_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemBytes);
		cudaFuncSetAttribute(hiz_tail_spread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemBytes);
		attrSet[dev & 63] = true;
	}
	if (p.exact_levels >= p.pyr.levels) return cudaSuccess;
	launch_tail(p, stream);
	return cudaGetLastError();
}
