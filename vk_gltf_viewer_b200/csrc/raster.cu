// raster.cu — software visibility-buffer rasteriser (sm_100a).
// Replaces, for every surviving MeshletDraw: shaders/visbuffer/visbuffer.mesh.glsl:30-104 (vertex transform +
// per-triangle facing cull), the fixed-function clip / viewport / raster / depth stages configured at
// application.cpp:326-340,772-841 (+ src/vulkan/pipeline_builder.cpp:225-277) and visbuffer.frag.glsl:36.
//
// One warp per meshlet, work-stealing over the survivor list in batches of kBatch:
//   * one lane per meshlet of the batch walks the dependent header chain (list -> draw -> primitive -> meshlet), so the
//     chain's L2 latency is paid once per batch, not once per meshlet;
//   * while meshlet j is processed, meshlet j+1's vertex indices are in flight in registers and its vertex positions,
//     triangle index words and mvp are copied global -> shared with cp.async (no registers held across the phases);
//   * mvp = viewProjection * transform and sign(det(transform)) come from a per-transform prologue kernel;
//   * vertex phase: clip position, outcodes and the snapped 24.8 screen position go to per-warp shared memory;
//   * triangle phase 1 (all triangles, one per lane): facing cull, trivial reject, pixel-centre bbox test; survivors are
//     compacted with ballot/popc into a shared list;
//   * triangle phase 2 (survivors only, full lanes): integer edge setup; small triangles are scanned by their lane with
//     int32 edge functions, large or clipped ones by the whole warp in 8x4 stamps with int64 edge functions;
//   * visibility: ONE fire-and-forget 64-bit RED.MIN per covered pixel on (~depthBits << 32 | drawId << 7 | triangle).
#include "kernels.cuh"
#include <cstdlib>

namespace {

#ifndef VKV_RASTER_WARPS
#define VKV_RASTER_WARPS 8
#endif
constexpr int kWarpsPerBlock = VKV_RASTER_WARPS;
constexpr int kThreads = kWarpsPerBlock * 32;
#ifndef VKV_SERIAL_MAX_DIM
#define VKV_SERIAL_MAX_DIM 8
#endif
constexpr int kSerialMaxDim = VKV_SERIAL_MAX_DIM;   // lane-serial path: bbox <= this many pixels in x and y
// Large triangles are not rasterised where they are found (one warp would scan thousands of stamps while the rest of the
// GPU idles): they go to a queue and a second kernel spreads their screen TILES over all warps.
#ifndef VKV_BIG_TILE_W
#define VKV_BIG_TILE_W 128
#endif
#ifndef VKV_BIG_TILE_H
#define VKV_BIG_TILE_H 64
#endif
constexpr int kBigTileW = VKV_BIG_TILE_W, kBigTileH = VKV_BIG_TILE_H;   // pixels per tile-work item (128 x 64 = 16 x 16 stamps of 8x4)
#ifndef VKV_BIG_MIN_STAMPS
#define VKV_BIG_MIN_STAMPS 1
#endif
constexpr int kBigMinStamps = VKV_BIG_MIN_STAMPS;  // bbox of >= this many 8x4 stamps -> deferred
#ifndef VKV_RASTER_BATCH
#define VKV_RASTER_BATCH 4      // meshlets per work-stealing grab on LONG lists (>= VKV_RASTER_LONG meshlets per resident warp): cfg 5 on one GPU
#endif                          // (1.5 M meshlets, 8K): grab 4 -> raster A 0.99 ms, grab 2 -> 1.11 ms (profiles/r4i) — the header chain and the copies
                                // of the next meshlet hide behind more work
#ifndef VKV_RASTER_BATCH_SHORT
#define VKV_RASTER_BATCH_SHORT 2 // ... and on ordinary lists.  A meshlet occupies a warp for ~8 us, so whole grabs leave the kernel's end ragged; measured
#endif                          // on cfg 3 (33 meshlets per warp; profiles/r4b): grab 1 -> 0.3763 ms, 2 -> 0.3806, 3 -> 0.3952, 4 -> 0.4089, 8 -> 0.4690
#ifndef VKV_RASTER_LONG
#define VKV_RASTER_LONG 64
#endif
#ifndef VKV_RASTER_GUIDED
#define VKV_RASTER_GUIDED 4     // guided self-scheduling: grabs shrink towards single meshlets once fewer than this many grabs per warp are left (0 = off;
#endif                          // profiles/r4c, cfg 3, grab 2: off -> 0.3806, 1 -> 0.3735, 2 -> 0.3682, 4 -> 0.3658)
#ifndef VKV_RASTER_MIN_BLOCKS
#define VKV_RASTER_MIN_BLOCKS 4
#endif
constexpr int kBatch = VKV_RASTER_BATCH;           // meshlets fetched per work-stealing grab
constexpr int kMinBlocks = VKV_RASTER_MIN_BLOCKS;  // blocks of 8 warps per SM the hot kernel is compiled for
constexpr int kDrainThreads = 256;

// what one lane fetches for one meshlet of a batch (mesh.glsl:31-36 resolved to addresses)
struct alignas(16) MeshletHdr {
	const uint32_t* vidx;      // primitive.vertexIndexBuffer + meshlet.vertexOffset
	const uint8_t* tri;        // primitive.primitiveIndexBuffer + meshlet.triangleOffset
	const void* verts;         // primitive.vertexBuffer (vkv_Vertex[]), or the 16-bit positions (int16 x, y, z, 0 per vertex) when bit 18 of counts is set
	uint32_t drawId;
	uint32_t tIdx;
	uint32_t counts;           // vertexCount | triangleCount << 8 | doubleSided << 16 | detNegative << 17 | quantized << 18 | normalized << 19
	uint32_t pad[3];
};

struct WarpScratch {
	float4 cxyw[VKV_MAX_VERTICES];   // clip x, y, w and the outcode bits (what phase 1 reads: one LDS.128 per vertex)
	int4 scr[VKV_MAX_VERTICES];      // snapped x, y (24.8), z_ndc bits, unused
	float cz[VKV_MAX_VERTICES];      // clip z (clipper only)
	float pos[VKV_MAX_VERTICES * 3]; // cp.async landing zone: object-space positions of the next meshlet, component k of vertices (lane, lane + 32)
	                                 // adjacent at [(k * 32 + lane) * 2 + half]: one 8-byte load gives the f32x2 pair the vertex phase works on
	uint32_t tri_words[2][96];       // cp.async landing zone (double buffered): up to 124*3 = 372 index bytes
	uint32_t surv[128];              // phase-1 survivors: ia | ib << 8 | ic << 16 | triangle << 24 | needsClip << 31
	float2 mvp2[4][4];               // cp.async landing zone: mvp[row][column], every element twice (f32x2 operand for two vertices)
	MeshletHdr hdr[kBatch];
#ifdef VKV_RASTER_BULK
	// bulk-async (TMA) landing zones: the 16-byte-aligned windows enclosing a meshlet's triangle bytes / vertex-index list, double buffered
	alignas(16) unsigned char tri_raw[2][416];   // <= 15 + 124 * 3 bytes, rounded up to 16
	alignas(16) unsigned char vidx_raw[2][288];  // <= 12 + 64 * 4 bytes, rounded up to 16
	alignas(8) unsigned long long mbar[4];       // completion barriers: tri_raw[0], tri_raw[1], vidx_raw[0], vidx_raw[1]
#endif
};
struct SlowScratch { Tri sub[8]; int nsub; }; // the overflow re-walk's clipper output (raster_big_kernel only)

enum { F_NEEDS_CLIP = 64, F_NAN = 128 };

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#ifdef VKV_RASTER_BULK
// cp.async.bulk (the TMA unit's linear copy): one elected lane moves a 16-byte-aligned window global -> shared and an mbarrier counts
// the bytes in; SASS: UBLKCP + SYNCS.  The vertex positions stay with LDGSTS: they are a gather through the index list.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic-proxy reads of the landing zone are ordered before the async write
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
	asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

__device__ __forceinline__ bool top_left(int dx, int dy) { return dy < 0 || (dy == 0 && dx > 0); }

__device__ __forceinline__ void project(float4 c, float hw, float hh, int& fx, int& fy, float& z) {
	// perspective divide: three IEEE divisions by the same w -> one refined reciprocal when every operand is a normal
	// number of moderate magnitude (bit-identical to `/`, see common.cuh div3_shared), the plain divisions otherwise
	float nx, ny;
	const float amx = max_nan(max_nan(fabsf(c.x), fabsf(c.y)), max_nan(fabsf(c.z), fabsf(c.w)));
	const float amn = fminf(fminf(fabsf(c.x), fabsf(c.y)), fminf(fabsf(c.z), fabsf(c.w)));
	if (amn >= kDivLo && amx <= kDivHi) div3_shared(c.x, c.y, c.z, c.w, nx, ny, z);
	else { nx = c.x / c.w; ny = c.y / c.w; z = c.z / c.w; }
	const float sx = nx * hw + hw;
	const float sy = ny * hh + hh;
	fx = __float2int_rn(sx * (float)VKV_SUB);
	fy = __float2int_rn(sy * (float)VKV_SUB);
}

// pixel-centre bounding box of three snapped vertices, clipped to the viewport; false = no pixel centre inside
__device__ __forceinline__ bool tri_bbox(int ax, int ay, int bx, int by, int cx, int cy, uint32_t W, uint32_t H, int& xmin, int& xmax,
                                         int& ymin, int& ymax, uint32_t& small) {
	const int minx = min(ax, min(bx, cx)), maxx = max(ax, max(bx, cx));
	const int miny = min(ay, min(by, cy)), maxy = max(ay, max(by, cy));
	xmin = max(0, (minx + (VKV_SUB / 2 - 1)) >> VKV_SUB_BITS);
	xmax = min((int)W - 1, (maxx - VKV_SUB / 2) >> VKV_SUB_BITS);
	ymin = max(0, (miny + (VKV_SUB / 2 - 1)) >> VKV_SUB_BITS);
	ymax = min((int)H - 1, (maxy - VKV_SUB / 2) >> VKV_SUB_BITS);
	small = (maxx - minx <= 16384 && maxy - miny <= 16384) ? 1u : 0u;
	return xmin <= xmax && ymin <= ymax;
}

// integer setup shared by every path; false = nothing to draw
__device__ __forceinline__ bool setup_tri(int ax, int ay, float za, int bx, int by, float zb, int cx, int cy, float zc, uint32_t id,
                                          uint32_t W, uint32_t H, Tri& t) {
	if (!tri_bbox(ax, ay, bx, by, cx, cy, W, H, t.xmin, t.xmax, t.ymin, t.ymax, t.small)) return false;
	long long area2;
	float fa;
	if (t.small) { // every delta < 2^15: both products < 2^29, the difference fits in int32 (same value as the int64 form)
		const int a32 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
		area2 = a32;
		fa = (float)abs(a32);
	} else {
		area2 = (long long)(bx - ax) * (cy - ay) - (long long)(by - ay) * (cx - ax);
		fa = (float)(area2 < 0 ? -area2 : area2);
	}
	if (area2 == 0) return false;
	if (area2 < 0) { // cullMode NONE: both windings are drawn
		int tx = bx; bx = cx; cx = tx;
		int ty = by; by = cy; cy = ty;
		float tz = zb; zb = zc; zc = tz;
		area2 = -area2;
	}
	t.ax = ax; t.ay = ay; t.bx = bx; t.by = by; t.cx = cx; t.cy = cy;
	t.area2 = area2;
	t.za = za; t.dzb = zb - za; t.dzc = zc - za;
	t.invA = 1.0f / fa;
	t.id = id;
	return true;
}

__device__ __forceinline__ void shade(unsigned long long* __restrict__ vis, uint32_t W, int x, int y, float w1, float w2, float invA,
                                      float za, float dzb, float dzc, uint32_t id) {
	const float l1 = w1 * invA;
	const float l2 = w2 * invA;
	float z = (za + l1 * dzb) + l2 * dzc;
	z = (z > 0.0f) ? z : 0.0f;
	z = (z < 1.0f) ? z : 1.0f;
	const unsigned long long key = ((unsigned long long)(~__float_as_uint(z)) << 32) | id;
	// result unused -> RED.E.MIN.64: fire and forget, the lane never waits for L2 (a read-before-atomic filter was measured
	// 30% slower on cfg 3: the L2 round trip per fragment costs more than the saved atomics)
	atomicMin(vis + (size_t)y * W + x, key);
}

// Lane-serial scan of a small triangle; all edge values fit in int32 (deltas <= 2^14 sub-pixels).
__device__ __forceinline__ void raster_serial(const Tri& t, unsigned long long* __restrict__ vis, uint32_t W) {
	const int e0dx = t.cx - t.bx, e0dy = t.cy - t.by;
	const int e1dx = t.ax - t.cx, e1dy = t.ay - t.cy;
	const int e2dx = t.bx - t.ax, e2dy = t.by - t.ay;
	const int b0 = top_left(e0dx, e0dy) ? 0 : 1, b1 = top_left(e1dx, e1dy) ? 0 : 1, b2 = top_left(e2dx, e2dy) ? 0 : 1;
	const int px0 = t.xmin * VKV_SUB + VKV_SUB / 2, py0 = t.ymin * VKV_SUB + VKV_SUB / 2;
	int r0 = e0dx * (py0 - t.by) - e0dy * (px0 - t.bx);
	int r1 = e1dx * (py0 - t.cy) - e1dy * (px0 - t.cx);
	int r2 = e2dx * (py0 - t.ay) - e2dy * (px0 - t.ax);
	for (int y = t.ymin; y <= t.ymax; ++y) {
		int w0 = r0, w1 = r1, w2 = r2;
		for (int x = t.xmin; x <= t.xmax; ++x) {
			if (w0 >= b0 && w1 >= b1 && w2 >= b2) shade(vis, W, x, y, (float)w1, (float)w2, t.invA, t.za, t.dzb, t.dzc, t.id);
			w0 -= e0dy * VKV_SUB; w1 -= e1dy * VKV_SUB; w2 -= e2dy * VKV_SUB;
		}
		r0 += e0dx * VKV_SUB; r1 += e1dx * VKV_SUB; r2 += e2dx * VKV_SUB;
	}
}

// Whole-warp scan of the pixel rectangle [x0,x1] x [y0,y1] (inside the triangle's bbox) in 8x4 stamps with stamp-level
// rejection; int64 edge functions evaluated from absolute pixel coordinates, so any partition of the bbox into rectangles
// produces exactly the same fragments.
__device__ __noinline__ void raster_coop(const Tri& t, unsigned long long* __restrict__ vis, uint32_t W, uint32_t lane, int x0, int x1, int y0, int y1) {
	const long long e0dx = t.cx - t.bx, e0dy = t.cy - t.by;
	const long long e1dx = t.ax - t.cx, e1dy = t.ay - t.cy;
	const long long e2dx = t.bx - t.ax, e2dy = t.by - t.ay;
	const long long b0 = top_left((int)e0dx, (int)e0dy) ? 0 : 1, b1 = top_left((int)e1dx, (int)e1dy) ? 0 : 1,
	                b2 = top_left((int)e2dx, (int)e2dy) ? 0 : 1;
	const int lx = lane & 7, ly = lane >> 3;
	// per-lane offset inside a stamp and the stamp-wide maximum of each edge function relative to its origin
	const long long o0 = e0dx * (ly * VKV_SUB) - e0dy * (lx * VKV_SUB);
	const long long o1 = e1dx * (ly * VKV_SUB) - e1dy * (lx * VKV_SUB);
	const long long o2 = e2dx * (ly * VKV_SUB) - e2dy * (lx * VKV_SUB);
	const long long m0 = max(0ll, e0dx * 3 * VKV_SUB) + max(0ll, -e0dy * 7 * VKV_SUB);
	const long long m1 = max(0ll, e1dx * 3 * VKV_SUB) + max(0ll, -e1dy * 7 * VKV_SUB);
	const long long m2 = max(0ll, e2dx * 3 * VKV_SUB) + max(0ll, -e2dy * 7 * VKV_SUB);
	const long long px0 = (long long)x0 * VKV_SUB + VKV_SUB / 2, py0 = (long long)y0 * VKV_SUB + VKV_SUB / 2;
	long long r0 = e0dx * (py0 - t.by) - e0dy * (px0 - t.bx);
	long long r1 = e1dx * (py0 - t.cy) - e1dy * (px0 - t.cx);
	long long r2 = e2dx * (py0 - t.ay) - e2dy * (px0 - t.ax);
	for (int ty = y0; ty <= y1; ty += 4) {
		long long s0 = r0, s1 = r1, s2 = r2;
		for (int tx = x0; tx <= x1; tx += 8) {
			if (s0 + m0 >= b0 && s1 + m1 >= b1 && s2 + m2 >= b2) {
				const long long w0 = s0 + o0, w1 = s1 + o1, w2 = s2 + o2;
				const int x = tx + lx, y = ty + ly;
				if (x <= x1 && y <= y1 && w0 >= b0 && w1 >= b1 && w2 >= b2)
					shade(vis, W, x, y, (float)w1, (float)w2, t.invA, t.za, t.dzb, t.dzc, t.id);
			}
			s0 -= e0dy * 8 * VKV_SUB; s1 -= e1dy * 8 * VKV_SUB; s2 -= e2dy * 8 * VKV_SUB;
		}
		r0 += e0dx * 4 * VKV_SUB; r1 += e1dx * 4 * VKV_SUB; r2 += e2dx * 4 * VKV_SUB;
	}
}

__device__ __forceinline__ bool is_big(const Tri& t) {
	return ((t.xmax - t.xmin + 8) >> 3) * ((t.ymax - t.ymin + 4) >> 2) >= kBigMinStamps;
}

// Append a large triangle to the queue (one lane).  ONE 64-bit atomic hands out the record slot (top 24 bits) and the
// triangle's range of tile-work indices (low 40 bits), so record order == tile-base order and the drain kernel can binary-search
// it.  40 bits cannot carry into the slot field: a triangle covers at most (32768/128) * (32768/64) = 2^17 tiles and fewer than
// 2^23 pushes can happen per launch beyond the capacity check below.  Stored records additionally keep their whole tile range
// below 2^32 (the drain kernel's work counter is 32 bits).  false = not queued: the caller rasterises the triangle in place.
constexpr unsigned long long kBigTileMask = (1ull << kBigSlotShift) - 1ull;
__device__ __forceinline__ bool push_big(const RasterParams& p, const Tri& t) {
	const uint32_t tilesX = (uint32_t)(t.xmax / kBigTileW - t.xmin / kBigTileW + 1), tilesY = (uint32_t)(t.ymax / kBigTileH - t.ymin / kBigTileH + 1);
	const uint32_t tiles = tilesX * tilesY;
	// full already (plain load; racy by at most one push per thread in flight, which the 24-bit slot field absorbs)
	if ((uint32_t)(*(volatile unsigned long long*)p.bigCursor >> kBigSlotShift) >= p.bigCap) return false;
	const unsigned long long old = atomicAdd(p.bigCursor, (1ull << kBigSlotShift) | (unsigned long long)tiles);
	const uint32_t slot = (uint32_t)(old >> kBigSlotShift);
	const unsigned long long base = old & kBigTileMask;
	// beyond the capacity, or the tile range would leave 32 bits: its range lies beyond every stored record's and is never visited
	if (slot >= p.bigCap) return false;
	if (base + tiles > 0xffffffffull) { // keeps its slot as an empty sentinel (such slots form a suffix: bases only grow)
		BigTri e = {};
		e.tileBase = 0xffffffffu;
		p.big[slot] = e;
		return false;
	}
	BigTri b;
	b.t = t; b.tileBase = (uint32_t)base; b.tilesX = tilesX; b.tilesY = tilesY; b.pad = 0;
	p.big[slot] = b;
	return true;
}

// Sutherland–Hodgman against one plane (dist >= 0 inside); intersections evaluated from the inside vertex outwards.
template <int PLANE>
__device__ __forceinline__ float plane_dist(const float4& v) {
	if (PLANE == 0) return v.w - v.z;               // near (reverse-Z: z <= w)
	if (PLANE == 1) return v.z;                     // far  (z >= 0)
	if (PLANE == 2) return VKV_GUARD * v.w - v.x;
	if (PLANE == 3) return VKV_GUARD * v.w + v.x;
	if (PLANE == 4) return VKV_GUARD * v.w - v.y;
	return VKV_GUARD * v.w + v.y;
}
template <int PLANE>
__device__ int clip_plane(const float4* in, int n, float4* out) {
	int m = 0;
	for (int i = 0; i < n; ++i) {
		const float4 cur = in[i];
		const float4 nxt = in[(i + 1 == n) ? 0 : i + 1];
		const float dc = plane_dist<PLANE>(cur), dn = plane_dist<PLANE>(nxt);
		const bool ic = dc >= 0.f, inx = dn >= 0.f;
		if (ic) out[m++] = cur;
		if (ic != inx) {
			const float4 a = ic ? cur : nxt, b = ic ? nxt : cur;
			const float da = ic ? dc : dn, db = ic ? dn : dc;
			const float t = da / (da - db);
			float4 r;
			r.x = a.x + t * (b.x - a.x);
			r.y = a.y + t * (b.y - a.y);
			r.z = a.z + t * (b.z - a.z);
			r.w = a.w + t * (b.w - a.w);
			out[m++] = r;
		}
	}
	return m;
}

// clip one triangle and emit up to 7 set-up fan triangles into `sub`
__device__ __noinline__ int clip_and_setup(float4 A, float4 B, float4 C, uint32_t id, uint32_t W, uint32_t H, Tri* sub) {
	float4 p0[12], p1[12];
	p0[0] = A; p0[1] = B; p0[2] = C;
	int n = 3;
	n = clip_plane<0>(p0, n, p1); if (n < 3) return 0;
	n = clip_plane<1>(p1, n, p0); if (n < 3) return 0;
	n = clip_plane<2>(p0, n, p1); if (n < 3) return 0;
	n = clip_plane<3>(p1, n, p0); if (n < 3) return 0;
	n = clip_plane<4>(p0, n, p1); if (n < 3) return 0;
	n = clip_plane<5>(p1, n, p0); if (n < 3) return 0;
	const float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
	int out = 0;
	for (int i = 1; i + 1 < n; ++i) {
		const float4 a = p0[0], b = p0[i], c = p0[i + 1];
		if (!(a.w > 0.f) || !(b.w > 0.f) || !(c.w > 0.f)) continue;
		int ax, ay, bx, by, cx, cy;
		float za, zb, zc;
		project(a, hw, hh, ax, ay, za);
		project(b, hw, hh, bx, by, zb);
		project(c, hw, hh, cx, cy, zc);
		if (setup_tri(ax, ay, za, bx, by, zb, cx, cy, zc, id, W, H, sub[out])) ++out;
	}
	return out;
}

// Per-transform prologue of the mesh shader (mesh.glsl:43-44,71), hoisted out of the per-meshlet path:
// mvp = viewProjection * transform (column by column) and the sign of determinant(transform).  Same arithmetic as before,
// computed once per mesh-node instead of once per meshlet.
__global__ void prepare_transforms_kernel(const float* __restrict__ transforms, const vkv_Camera* __restrict__ camera, uint32_t n,
                                          float* __restrict__ mvpOut, uint32_t* __restrict__ detNeg, float4* __restrict__ eye) {
	__shared__ float sVP[16];
	if (threadIdx.x < 16) sVP[threadIdx.x] = __ldg(camera->viewProjection + threadIdx.x);
	__syncthreads();
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
		transform_prologue(transforms + (size_t)t * 16, sVP, mvpOut + (size_t)t * 16, detNeg + t, eye ? eye + t : nullptr);
}

// Strip mode (strips.cu): every 64x16-pixel tile a drawn triangle's pixel bounding box touches gets its dirty byte set, so that
// the strip owners pull only tiles somebody drew into.  Conservative (the box, not the coverage) and idempotent.
// Strip mode (every pass marks, millions of triangles hit the same few bytes): a flag that already reads 1 is not stored again — the few KB
// of flags live in L1, the stores go to L2 (measured on a 54 k-meshlet pass: unconditional stores 15 us, checked ones 5 us of an 89 us
// launch); a stale 0 (another SM set the byte) only costs a redundant store.  Small pass-B launches store unconditionally: the load
// would sit on the triangle's critical path (cfg 3 pass B: 36.0 us against 39.5).
__device__ __forceinline__ void mark_tile(uint8_t* q, bool check) { if (!check || __ldca(q) == 0) *q = 1; }
// Are tiles marked in this launch?  Strip mode: always (markLimit = ~0).  Pass-B pyramid rebuild: only while the pass is small — the marks
// of a large pass B cost more than the pyramid kernel saves (and dirty nearly every tile anyway); the pyramid kernel reads the same
// counter and takes the same decision (HizParams::dirty_count / dirty_limit).
__device__ __forceinline__ bool marking_on(const RasterParams& p) { return p.dirty != nullptr && __ldg(p.count) <= p.markLimit; }
__device__ __forceinline__ void mark_small(const RasterParams& p, const Tri& t) { // bbox <= 8x8 pixels: at most 2x2 tiles
	const uint32_t tx0 = (uint32_t)t.xmin >> 6, tx1 = (uint32_t)t.xmax >> 6, ty0 = (uint32_t)t.ymin >> 4, ty1 = (uint32_t)t.ymax >> 4;
	uint8_t* d = p.dirty + ty0 * p.dirtyTilesX;
	const bool check = p.markLimit == 0xffffffffu;
	mark_tile(d + tx0, check);
	if (tx1 != tx0) mark_tile(d + tx1, check);
	if (ty1 != ty0) {
		d += p.dirtyTilesX;
		mark_tile(d + tx0, check);
		if (tx1 != tx0) mark_tile(d + tx1, check);
	}
}
__device__ __forceinline__ void mark_rect(const RasterParams& p, bool marking, int x0, int x1, int y0, int y1, uint32_t lane) { // whole warp
	if (!marking || x0 > x1 || y0 > y1) return;
	const uint32_t tx0 = (uint32_t)x0 >> 6, tx1 = (uint32_t)x1 >> 6, ty0 = (uint32_t)y0 >> 4, ty1 = (uint32_t)y1 >> 4;
	for (uint32_t ty = ty0; ty <= ty1; ++ty)
		for (uint32_t tx = tx0 + lane; tx <= tx1; tx += 32) mark_tile(p.dirty + ty * p.dirtyTilesX + tx, true);
}

// Append a triangle that needs the clipper to the clip queue (one lane).  false = queue full.
__device__ __forceinline__ bool push_clip(const RasterParams& p, const float4& A, const float4& B, const float4& C, uint32_t id) {
	if (*(volatile uint32_t*)p.clipCount >= p.clipCap) return false; // keeps the counter from running away once full
	const uint32_t slot = atomicAdd(p.clipCount, 1u);
	if (slot >= p.clipCap) return false;
	ClipTri c;
	c.a = A; c.b = B; c.c = C; c.id = id; c.pad[0] = c.pad[1] = c.pad[2] = 0;
	p.clip[slot] = c;
	return true;
}

// The meshlet loop, in two instantiations of the same text:
//   kHot = true   raster_kernel.  Everything a lane can finish alone — vertex transform, facing cull, set-up, the lane-serial
//                 scan of small triangles.  What it cannot (triangles that need the clipper, triangles larger than the serial
//                 limit) is QUEUED for raster_big_kernel, so neither the clipper's stack arrays nor the cooperative scan's
//                 64-bit edge functions count against this kernel's registers: 3 blocks of 8 warps per SM instead of 2.
//   kHot = false  the overflow re-walk inside raster_big_kernel (only when a queue was full): the same loop, skipping what the
//                 hot kernel already drew and clipping / scanning everything else in place.  A triangle drawn twice is harmless:
//                 the visibility write is an atomic min.
template <bool kHot>
__device__ __forceinline__ void meshlet_loop(const RasterParams& p, WarpScratch& ws, SlowScratch* slow, uint32_t* __restrict__ workCursor, uint32_t lane) {
	const uint32_t count = __ldg(p.count);
	const bool marking = p.dirty != nullptr && count <= p.markLimit;
	const float hw = (float)p.W * 0.5f, hh = (float)p.H * 0.5f;
	const uint32_t below = (1u << lane) - 1u;
#ifdef VKV_RASTER_BULK
	if (lane == 0) {
		for (int b = 0; b < 4; ++b) mbar_init(&ws.mbar[b]);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	uint32_t phases = 0; // bit b: the parity mbar[b]'s next completion will have
#endif

	// meshlets per work-stealing grab: kBatch when there is plenty of work (the header chain's latency is paid once per batch);
	// fewer when the list is short, so that a small scene spreads over all warps instead of queueing 4 deep behind a few
	const uint32_t totalWarps = gridDim.x * (blockDim.x >> 5);
	const uint32_t perWarp = count / totalWarps;
	const uint32_t batch0 = !kHot ? (uint32_t)kBatch : perWarp >= (uint32_t)VKV_RASTER_LONG ? (uint32_t)kBatch : min((uint32_t)min(kBatch, VKV_RASTER_BATCH_SHORT), max(1u, perWarp));
	uint32_t batch = batch0;
#if VKV_RASTER_GUIDED
	uint32_t seen = 0; // where the cursor stood at this warp's last grab
#endif
	for (;;) {
#if VKV_RASTER_GUIDED
		// guided self-scheduling: a meshlet keeps a warp busy for microseconds, so a kernel that ends on whole batches ends ragged (the
		// last warps still hold batch - 1 meshlets when the others run dry).  Grabs shrink as the list runs out: all warps have grabbed
		// about once since `seen`, so about count - seen - totalWarps * batch meshlets are left; spread them VKV_RASTER_GUIDED grabs deep.
		if (kHot) {
			const uint32_t taken = seen + totalWarps * batch;
			const uint32_t left = count > taken ? count - taken : 0u;
			batch = min(batch0, max(1u, left / (totalWarps * (uint32_t)VKV_RASTER_GUIDED)));
		}
#endif
		uint32_t base = 0;
		if (lane == 0) base = atomicAdd(workCursor, batch);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= count) break;
#if VKV_RASTER_GUIDED
		seen = base;
#endif
		const uint32_t nb = min(batch, count - base);
		// one lane per meshlet walks the dependent chain list -> draw -> primitive -> meshlet (mesh.glsl:31-36)
		if (lane < nb) {
			const uint32_t drawId = __ldg(p.list + base + lane);
			const vkv_MeshletDraw* d = p.draws + drawId;
			const uint32_t primIdx = __ldg(&d->primitiveIndex), mlIdx = __ldg(&d->meshletIndex), tIdx = __ldg(&d->transformIndex);
			const vkv_Primitive* prim = p.primitives + primIdx;
			const ulonglong2 b0 = __ldg((const ulonglong2*)prim);           // vertexIndexBuffer, primitiveIndexBuffer
			const ulonglong2 b1 = __ldg((const ulonglong2*)prim + 1);       // vertexBuffer, meshletBuffer
			const uint32_t matIdx = __ldg(&prim->materialIndex);
			const vkv_Meshlet* ml = (const vkv_Meshlet*)b1.y + mlIdx;
			const uint32_t vertexOffset = __ldg(&ml->vertexOffset), triangleOffset = __ldg(&ml->triangleOffset);
			const uint32_t cnt = __ldg((const uint32_t*)&ml->vertexCount);
			const uint32_t vc = min(cnt & 0xffu, VKV_MAX_VERTICES), tc = min((cnt >> 8) & 0xffu, VKV_MAX_MESHLET_TRIANGLES);
			const uint32_t ds = __ldg(&p.materials[matIdx].doubleSided) != 0 ? 1u : 0u;
			const uint32_t dn = __ldg(p.detNeg + tIdx);
			MeshletHdr h;
			h.vidx = (const uint32_t*)b0.x + vertexOffset;
			h.tri = (const uint8_t*)b0.y + triangleOffset;
			h.verts = (const void*)b1.x;
			h.drawId = drawId; h.tIdx = tIdx;
			h.counts = vc | (tc << 8) | (ds << 16) | (dn << 17);
			if (p.qtable) { // KHR_mesh_quantization kept in its 16-bit form: 8 bytes per vertex instead of a 24-byte record
				const ulonglong2 q = __ldg((const ulonglong2*)(p.qtable + primIdx)); // positions | normalized, reserved
				if (q.x) { h.verts = (const void*)q.x; h.counts |= (1u << 18) | (((uint32_t)q.y & 1u) << 19); }
			}
			ws.hdr[lane] = h;
		}
		__syncwarp();

		uint32_t vi0 = 0, vi1 = 0; // vertex indices (slots lane, lane+32) of the meshlet whose copies are issued next
#ifdef VKV_RASTER_BULK
		// the index list as ONE bulk copy issued by lane 0 into vidx_raw[buf]; fetch_indices() reads it back once its barrier completes
		auto load_indices = [&](const MeshletHdr& h, uint32_t buf) {
			if (lane == 0) {
				const uintptr_t a = (uintptr_t)h.vidx;
				const uint32_t off = (uint32_t)(a & 15), bytes = (off + (h.counts & 0xffu) * 4u + 15u) & ~15u;
				bulk_load(ws.vidx_raw[buf], (const void*)(a - off), bytes, &ws.mbar[2 + buf]);
			}
		};
		auto fetch_indices = [&](const MeshletHdr& h, uint32_t buf) {
			mbar_wait(&ws.mbar[2 + buf], (phases >> (2 + buf)) & 1u);
			phases ^= 1u << (2 + buf);
			const uint32_t vc = h.counts & 0xffu;
			const uint32_t* v = (const uint32_t*)(ws.vidx_raw[buf] + ((uintptr_t)h.vidx & 15));
			if (lane < vc) vi0 = v[lane];
			if (lane + 32 < vc) vi1 = v[lane + 32];
		};
#else
		auto load_indices = [&](const MeshletHdr& h, uint32_t) {
			const uint32_t vc = h.counts & 0xffu;
			if (lane < vc) vi0 = __ldg(h.vidx + lane);
			if (lane + 32 < vc) vi1 = __ldg(h.vidx + lane + 32);
		};
		auto fetch_indices = [&](const MeshletHdr&, uint32_t) {};
#endif
		// positions (vi0/vi1 hold the meshlet's indices), triangle index words and mvp: global -> shared, asynchronously
		auto issue_copies = [&](const MeshletHdr& h, uint32_t buf) {
			const uint32_t vc = h.counts & 0xffu, tc = (h.counts >> 8) & 0xffu;
			fetch_indices(h, buf);
			if (h.counts & (1u << 18)) { // 16-bit positions: one 8-byte copy per vertex into the same landing zone, vertex v at bytes [8v, 8v + 8)
				const uint2* qv = (const uint2*)h.verts;
				if (lane < vc) cp_async8((uint2*)ws.pos + lane, qv + vi0);
				if (lane + 32 < vc) cp_async8((uint2*)ws.pos + 32 + lane, qv + vi1);
			} else {
				const vkv_Vertex* fv = (const vkv_Vertex*)h.verts;
				if (lane < vc) {
					const float* q = fv[vi0].position;
					cp_async4(&ws.pos[lane * 2], q); cp_async4(&ws.pos[(32 + lane) * 2], q + 1); cp_async4(&ws.pos[(64 + lane) * 2], q + 2);
				}
				if (lane + 32 < vc) {
					const float* q = fv[vi1].position;
					cp_async4(&ws.pos[lane * 2 + 1], q); cp_async4(&ws.pos[(32 + lane) * 2 + 1], q + 1); cp_async4(&ws.pos[(64 + lane) * 2 + 1], q + 2);
				}
			}
#ifdef VKV_RASTER_BULK
			if (lane == 0) { // the triangle bytes as ONE bulk copy of their enclosing 16-byte-aligned window (any alignment of the slice)
				const uintptr_t a = (uintptr_t)h.tri;
				const uint32_t off = (uint32_t)(a & 15), bytes = (off + tc * 3u + 15u) & ~15u;
				bulk_load(ws.tri_raw[buf], (const void*)(a - off), bytes, &ws.mbar[buf]);
			}
			const uint32_t nWords = 0;
			if (false) {
#else
			const uint32_t nWords = (tc * 3 + 3) >> 2;
			if ((((uintptr_t)h.tri) & 3) == 0) {
#endif
				const uint32_t* w = (const uint32_t*)h.tri;
#pragma unroll
				for (uint32_t k = 0; k < 96; k += 32)
					if (lane + k < nWords) cp_async4(&ws.tri_words[buf][lane + k], w + lane + k);
			} else { // unaligned triangle slice (never produced by the reference's builder: assets.cpp:339 pads to 4)
				const uint32_t nBytes = tc * 3;
				for (uint32_t wi = lane; wi < nWords; wi += 32) {
					uint32_t r = 0;
					for (uint32_t b = 0; b < 4; ++b)
						if (wi * 4 + b < nBytes) r |= (uint32_t)__ldg(h.tri + wi * 4 + b) << (8 * b);
					ws.tri_words[buf][wi] = r;
				}
			}
			{ // mvp element e = lane >> 1 (column e >> 2, row e & 3), copied twice: lanes 2e and 2e + 1
				const uint32_t e = lane >> 1;
				cp_async4(&ws.mvp2[e & 3][e >> 2].x + (lane & 1), p.mvp + (size_t)h.tIdx * 16 + e);
			}
			cp_async_commit();
		};
		load_indices(ws.hdr[0], 0);
		issue_copies(ws.hdr[0], 0);

		for (uint32_t j = 0; j < nb; ++j) {
			const MeshletHdr h = ws.hdr[j];
			const uint32_t drawId = h.drawId;
			const uint32_t vc = h.counts & 0xffu, tc = (h.counts >> 8) & 0xffu;
			const bool doubleSided = (h.counts >> 16) & 1u, detNeg = (h.counts >> 17) & 1u;
			// vertex indices of meshlet j+1: in flight during this meshlet's vertex phase
			if (j + 1 < nb) load_indices(ws.hdr[j + 1], (j + 1) & 1);
			cp_async_wait_all();
#ifdef VKV_RASTER_BULK
			mbar_wait(&ws.mbar[j & 1], (phases >> (j & 1)) & 1u);
			phases ^= 1u << (j & 1);
#endif
			__syncwarp();

			// :50-69 vertices.  A lane owns vertices `lane` and `lane + 32`; their transform, perspective divide and viewport
			// mapping run as the two halves of packed f32x2 instructions (every half an individually rounded IEEE operation: the
			// same values as the scalar form, half the issue slots — see common.cuh / cull.cu for the exactness argument).
			{
				const f2 nz = p.neg_zero2;
				f2 P0, P1, P2;
				if (h.counts & (1u << 18)) {
					// dequantise in registers: fastgltf's convertComponent<float, int16_t> (tools.hpp:266-289) — float(x), or
					// max(float(x) / 32767, -1) for a normalized accessor — the very floats the host expansion writes into Vertex.position
					const uint2 a = ((const uint2*)ws.pos)[lane], b = ((const uint2*)ws.pos)[32 + lane];
					float ax = (float)(short)(a.x & 0xffffu), ay = (float)(short)(a.x >> 16), az = (float)(short)(a.y & 0xffffu);
					float bx = (float)(short)(b.x & 0xffffu), by = (float)(short)(b.x >> 16), bz = (float)(short)(b.y & 0xffffu);
					if (h.counts & (1u << 19)) {
						ax = fmaxf(__fdiv_rn(ax, 32767.0f), -1.0f); ay = fmaxf(__fdiv_rn(ay, 32767.0f), -1.0f); az = fmaxf(__fdiv_rn(az, 32767.0f), -1.0f);
						bx = fmaxf(__fdiv_rn(bx, 32767.0f), -1.0f); by = fmaxf(__fdiv_rn(by, 32767.0f), -1.0f); bz = fmaxf(__fdiv_rn(bz, 32767.0f), -1.0f);
					}
					P0 = pk(ax, bx); P1 = pk(ay, by); P2 = pk(az, bz);
				} else {
					P0 = *(const f2*)&ws.pos[lane * 2]; P1 = *(const f2*)&ws.pos[(32 + lane) * 2]; P2 = *(const f2*)&ws.pos[(64 + lane) * 2];
				}
				f2 C[4]; // clip x, y, z, w of both vertices: ((c0*x + c1*y) + c2*z) + c3   (:61, w = 1)
#pragma unroll
				for (int r = 0; r < 4; ++r) {
					const ulonglong2 m01 = *(const ulonglong2*)&ws.mvp2[r][0], m23 = *(const ulonglong2*)&ws.mvp2[r][2];
					C[r] = add2(add2(add2(mul2(m01.x, P0, nz), mul2(m01.y, P1, nz)), mul2(m23.x, P2, nz)), m23.y);
				}
				// perspective divide + viewport for both at once where both are in the shared-reciprocal range (common.cuh)
				const f2 NW = mul2(C[3], pk(-1.0f, -1.0f), nz);
				const f2 R = refined_rcp2(NW, pk(1.0f, 1.0f));
				const f2 NX = div_by2(C[0], NW, R, nz), NY = div_by2(C[1], NW, R, nz), NZ = div_by2(C[2], NW, R, nz);
				const f2 HW = pk(hw, hw), HH = pk(hh, hh), SUBP = pk((float)VKV_SUB, (float)VKV_SUB);
				const f2 FX = mul2(add2(mul2(NX, HW, nz), HW), SUBP, nz), FY = mul2(add2(mul2(NY, HH, nz), HH), SUBP, nz);
#pragma unroll
				for (int half = 0; half < 2; ++half) {
					const uint32_t v = lane + half * 32;
					if (v < vc) {
						const float4 c = half ? make_float4(hi_of(C[0]), hi_of(C[1]), hi_of(C[2]), hi_of(C[3])) : make_float4(lo_of(C[0]), lo_of(C[1]), lo_of(C[2]), lo_of(C[3]));
						uint32_t f = 0;
						if (c.x < -c.w) f |= 1;
						if (c.x > c.w) f |= 2;
						if (c.y < -c.w) f |= 4;
						if (c.y > c.w) f |= 8;
						if (c.z < 0.f) f |= 16;
						if (c.z > c.w) f |= 32;
						const float g = VKV_GUARD * c.w;
						if (c.z < 0.f || c.z > c.w || c.x > g || c.x < -g || c.y > g || c.y < -g) f |= F_NEEDS_CLIP;
						if (!(c.x == c.x && c.y == c.y && c.z == c.z && c.w == c.w)) f |= F_NAN; // NaN -> reject
						int fx = 0, fy = 0;
						float z = 0.f;
						if (!(f & (F_NEEDS_CLIP | F_NAN)) && c.w > 0.f) {
							const float amx = max_nan(max_nan(fabsf(c.x), fabsf(c.y)), max_nan(fabsf(c.z), fabsf(c.w)));
							const float amn = fminf(fminf(fabsf(c.x), fabsf(c.y)), fminf(fabsf(c.z), fabsf(c.w)));
							if (amn >= kDivLo && amx <= kDivHi) { // the packed results are the IEEE quotients (vkv_selftest_division)
								fx = __float2int_rn(half ? hi_of(FX) : lo_of(FX));
								fy = __float2int_rn(half ? hi_of(FY) : lo_of(FY));
								z = half ? hi_of(NZ) : lo_of(NZ);
							} else project(c, hw, hh, fx, fy, z);
						} else if (!(f & F_NAN) && !(c.w > 0.f)) f |= F_NEEDS_CLIP; // degenerate w: let the clipper decide
						ws.cxyw[v] = make_float4(c.x, c.y, c.w, __uint_as_float(f));
						ws.scr[v] = make_int4(fx, fy, __float_as_int(z), 0);
						ws.cz[v] = c.z;
					}
				}
			}
			__syncwarp();
			// positions / triangle words / mvp of meshlet j+1: in flight during this meshlet's triangle phases
			if (j + 1 < nb) issue_copies(ws.hdr[j + 1], (j + 1) & 1);

			// :73-103 triangles, phase 1: facing cull + trivial reject + bbox test, survivors compacted
#ifdef VKV_RASTER_BULK
			const uint8_t* tb = ws.tri_raw[j & 1] + ((uintptr_t)h.tri & 15);
#else
			const uint8_t* tb = (const uint8_t*)ws.tri_words[j & 1];
#endif
			uint32_t nSurv = 0;
			for (uint32_t tbase = 0; tbase < tc; tbase += 32) {
				const uint32_t t = tbase + lane;
				bool keep = false;
				uint32_t entry = 0;
				if (t < tc) {
					uint32_t ia = tb[t * 3], ib = tb[t * 3 + 1], ic = tb[t * 3 + 2];
					ia = min(ia, vc - 1); ib = min(ib, vc - 1); ic = min(ic, vc - 1); // robustness only
					const float4 A = ws.cxyw[ia], B = ws.cxyw[ib], C = ws.cxyw[ic];  // x, y, w, outcode
					const uint32_t fa = __float_as_uint(A.w), fb = __float_as_uint(B.w), fc = __float_as_uint(C.w);
					bool cull = false;
					if (!doubleSided) { // :86-98
						const float det = det3(make_float3(A.x, A.y, A.z), make_float3(B.x, B.y, B.z), make_float3(C.x, C.y, C.z));
						cull = detNeg ? (det < 0.0f) : (det > 0.0f);
					}
					if (!cull && !((fa | fb | fc) & F_NAN) && !(fa & fb & fc & 63)) {
						entry = ia | (ib << 8) | (ic << 16) | (t << 24);
						if ((fa | fb | fc) & F_NEEDS_CLIP) { keep = true; entry |= 0x80000000u; }
						else {
							const int4 a = ws.scr[ia], b = ws.scr[ib], c = ws.scr[ic];
							int x0, x1, y0, y1;
							uint32_t sm;
							keep = tri_bbox(a.x, a.y, b.x, b.y, c.x, c.y, p.W, p.H, x0, x1, y0, y1, sm);
						}
					}
				}
				const uint32_t m = __ballot_sync(0xffffffffu, keep);
				if (keep) ws.surv[nSurv + __popc(m & below)] = entry;
				nSurv += __popc(m);
			}
			__syncwarp();

			// phase 2: survivors only — edge setup and rasterisation
			for (uint32_t sbase = 0; sbase < nSurv; sbase += 32) {
				const uint32_t s = sbase + lane;
				int kind = 0; // 0 nothing, 1 lane-serial, 2 larger than the serial limit, 3 needs the clipper
				Tri tri;
				uint32_t entry = 0, id = 0;
				if (s < nSurv) {
					entry = ws.surv[s];
					const uint32_t ia = entry & 0xffu, ib = (entry >> 8) & 0xffu, ic = (entry >> 16) & 0xffu;
					id = (drawId << VKV_TRIANGLE_BITS) | ((entry >> 24) & 0x7fu); // frag.glsl:36
					if (entry & 0x80000000u) kind = 3;
					else {
						const int4 a = ws.scr[ia], b = ws.scr[ib], c = ws.scr[ic];
						if (setup_tri(a.x, a.y, __int_as_float(a.z), b.x, b.y, __int_as_float(b.z), c.x, c.y, __int_as_float(c.z), id, p.W, p.H, tri))
							kind = (tri.small && tri.xmax - tri.xmin < kSerialMaxDim && tri.ymax - tri.ymin < kSerialMaxDim) ? 1 : 2;
					}
				}
				if (kHot) {
					if (kind == 1) {
						if (marking) mark_small(p, tri);
						raster_serial(tri, p.vis, p.W);
					} else if (kind == 2) { if (!push_big(p, tri)) *p.overflow = 1u; }
					else if (kind == 3) {
						const uint32_t ia = entry & 0xffu, ib = (entry >> 8) & 0xffu, ic = (entry >> 16) & 0xffu;
						const float4 A = ws.cxyw[ia], B = ws.cxyw[ib], C = ws.cxyw[ic];
						if (!push_clip(p, make_float4(A.x, A.y, ws.cz[ia], A.z), make_float4(B.x, B.y, ws.cz[ib], B.z), make_float4(C.x, C.y, ws.cz[ic], C.z), id))
							*p.overflow = 1u;
					}
				} else { // overflow re-walk: everything the hot kernel did NOT draw itself, in place, one triangle at a time
					uint32_t todo = __ballot_sync(0xffffffffu, kind >= 2);
					while (todo) {
						const int src = __ffs(todo) - 1;
						todo &= todo - 1;
						if ((int)lane == src) {
							if (kind == 2) { slow->sub[0] = tri; slow->nsub = 1; }
							else {
								const uint32_t ia = entry & 0xffu, ib = (entry >> 8) & 0xffu, ic = (entry >> 16) & 0xffu;
								const float4 A = ws.cxyw[ia], B = ws.cxyw[ib], C = ws.cxyw[ic];
								slow->nsub = clip_and_setup(make_float4(A.x, A.y, ws.cz[ia], A.z), make_float4(B.x, B.y, ws.cz[ib], B.z),
								                            make_float4(C.x, C.y, ws.cz[ic], C.z), id, p.W, p.H, slow->sub);
							}
						}
						__syncwarp();
						const int n = slow->nsub;
						for (int k = 0; k < n; ++k) {
							mark_rect(p, marking, slow->sub[k].xmin, slow->sub[k].xmax, slow->sub[k].ymin, slow->sub[k].ymax, lane);
							raster_coop(slow->sub[k], p.vis, p.W, lane, slow->sub[k].xmin, slow->sub[k].xmax, slow->sub[k].ymin, slow->sub[k].ymax);
						}
						__syncwarp();
					}
				}
			}
			__syncwarp();
		}
	}
}

// out of line: the re-walk's register appetite (and, under the drain kernel's register cap, its spills) stays inside this function
__device__ __noinline__ void rewalk(const RasterParams& p, WarpScratch& ws, SlowScratch* slow, uint32_t lane) {
	meshlet_loop<false>(p, ws, slow, p.slowWork, lane);
}

__global__ void __launch_bounds__(kThreads, kMinBlocks) raster_kernel(const RasterParams p) {
	__shared__ WarpScratch scratch[kWarpsPerBlock];
	asm volatile("griddepcontrol.wait;" ::: "memory"); // vkv_frame launches this kernel right behind the cull: the survivor list is complete from here on
	meshlet_loop<true>(p, scratch[threadIdx.x >> 5], nullptr, p.work, threadIdx.x & 31);
	// Programmatic dependent launch: this warp is out of work.  Once every block has said so (or exited), the drain kernel behind
	// this one — launched with programmatic stream serialization — may have its blocks placed on the SMs that fall idle during
	// this kernel's tail; they wait in griddepcontrol.wait until the whole grid has completed and its writes are visible.
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Grid barrier of the drain kernel (all its blocks are co-resident: the launch sizes the grid from the occupancy query).
__device__ __forceinline__ void drain_grid_barrier(uint32_t* counter) {
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		atomicAdd(counter, 1u);
		while (*(volatile uint32_t*)counter < gridDim.x) __nanosleep(64);
		__threadfence();
	}
	__syncthreads();
}

// Drain of the two queues raster_kernel leaves behind:
//   1. clip queue: one warp per triangle — one lane clips and sets the pieces up; a piece spanning several 128x64 tiles joins
//      the large-triangle queue, the others are scanned by the warp at once.  (Skipped, barrier included, when the queue is empty.)
//   2. large-triangle queue: one warp per (triangle, 128x64-pixel tile) work item.
//   3. only if a queue overflowed: the re-walk of the meshlet list (meshlet_loop<false>), one warp per block.
#ifndef VKV_DRAIN_MINB
#define VKV_DRAIN_MINB 3
#endif
__global__ void __launch_bounds__(kDrainThreads, VKV_DRAIN_MINB) raster_big_kernel(const RasterParams p) {
	__shared__ Tri sTri[kDrainThreads / 32];
	__shared__ Tri sSub[kDrainThreads / 32][8];
#ifndef VKV_DRAIN_UNBATCHED
	__shared__ Tri sTriB[kDrainThreads / 32][32];
#endif
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); // the pyramid build behind this kernel may set its blocks up early too
	asm volatile("griddepcontrol.wait;" ::: "memory");              // raster_kernel has completed, its queue and its visbuffer writes are visible
	// the three words that decide what there is to do, fetched together (one L2 round trip, not three): all are final when this
	// kernel starts, except the queue cursor when the clip phase below appends pieces (re-read behind its barrier)
	const uint32_t clipCountNow = *(volatile uint32_t*)p.clipCount;
	unsigned long long cur = *(volatile unsigned long long*)p.bigCursor;
	const uint32_t overflowed = *(volatile uint32_t*)p.overflow;
	const bool marking = marking_on(p);
	if (p.drainSeen && blockIdx.x == 0 && threadIdx.x == 0) *p.drainSeen = clipCountNow + (uint32_t)(cur >> kBigSlotShift) + overflowed;
	const uint32_t nClip = min(clipCountNow, p.clipCap);
	if (nClip) {
		__shared__ int sN[kDrainThreads / 32];
		for (;;) {
			uint32_t ci = 0;
			if (lane == 0) ci = atomicAdd(p.clipNext, 1u);
			ci = __shfl_sync(0xffffffffu, ci, 0);
			if (ci >= nClip) break;
			__syncwarp();
			if (lane == 0) {
				const ClipTri c = p.clip[ci];
				const int n = clip_and_setup(c.a, c.b, c.c, c.id, p.W, p.H, sSub[warp]);
				int kept = 0; // clipped pieces are often the largest triangles of a scene: spread those over the GPU too
				for (int k = 0; k < n; ++k) {
					const Tri& t = sSub[warp][k];
					const bool multiTile = (t.xmax / kBigTileW != t.xmin / kBigTileW) || (t.ymax / kBigTileH != t.ymin / kBigTileH);
					if (multiTile && push_big(p, t)) continue;
					if (kept != k) sSub[warp][kept] = t;
					++kept;
				}
				sN[warp] = kept;
			}
			__syncwarp();
			const int n = sN[warp];
			for (int k = 0; k < n; ++k) {
				const Tri& t = sSub[warp][k];
				mark_rect(p, marking, t.xmin, t.xmax, t.ymin, t.ymax, lane);
				raster_coop(t, p.vis, p.W, lane, t.xmin, t.xmax, t.ymin, t.ymax);
			}
		}
		drain_grid_barrier(p.drainBarrier); // every piece has been queued before anyone reads the queue's extent
		cur = *(volatile unsigned long long*)p.bigCursor;
	}

	uint32_t nRec = min((uint32_t)(cur >> kBigSlotShift), p.bigCap);
	if ((cur & kBigTileMask) > 0xffffffffull) {
		uint32_t lo = 0, hi = nRec; // first sentinel slot
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if (p.big[mid].tilesX == 0) hi = mid; else lo = mid + 1;
		}
		nRec = lo;
	}
	if (nRec) {
		const BigTri* last = p.big + (nRec - 1);
		const uint32_t nTiles = last->tileBase + last->tilesX * last->tilesY;
#ifndef VKV_DRAIN_UNBATCHED
		// A warp claims up to 32 work items with ONE atomic; every lane finds the record of its own item (32 binary searches side by
		// side) and stages the triangle in shared memory; then the warp scans the items one after the other with nothing but the
		// visibility atomics on the memory path.  A scene full of medium triangles (one item each) is bound by exactly these
		// round trips; a scene of a few huge triangles keeps the claim small so that the last tiles still spread over the GPU.
		const uint32_t totalWarps = gridDim.x * (kDrainThreads / 32);
		const uint32_t claim = min(32u, max(1u, nTiles / (totalWarps * 4u)));
		for (;;) {
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(p.bigNext, claim);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= nTiles) break;
			const uint32_t n = min(claim, nTiles - base);
			uint32_t local = 0, tilesX = 1;
			__syncwarp();
			if (lane < n) {
				const uint32_t w = base + lane;
				uint32_t lo = 0, hi = nRec - 1; // last record with tileBase <= w
				while (lo < hi) {
					const uint32_t mid = (lo + hi + 1) >> 1;
					if (__ldg(&p.big[mid].tileBase) <= w) lo = mid; else hi = mid - 1;
				}
				const BigTri* b = p.big + lo;
				const unsigned long long* src = (const unsigned long long*)&b->t;
				unsigned long long* dst = (unsigned long long*)&sTriB[warp][lane];
#pragma unroll
				for (int k = 0; k < (int)(sizeof(Tri) / 8); ++k) dst[k] = __ldg(src + k);
				local = w - __ldg(&b->tileBase);
				tilesX = __ldg(&b->tilesX);
			}
			__syncwarp();
			for (uint32_t i = 0; i < n; ++i) {
				const Tri& t = sTriB[warp][i];
				const uint32_t li = __shfl_sync(0xffffffffu, local, i), txs = __shfl_sync(0xffffffffu, tilesX, i);
				const int tx = t.xmin / kBigTileW + (int)(li % txs), ty = t.ymin / kBigTileH + (int)(li / txs);
				const int x0 = max(t.xmin, tx * kBigTileW), x1 = min(t.xmax, tx * kBigTileW + kBigTileW - 1);
				const int y0 = max(t.ymin, ty * kBigTileH), y1 = min(t.ymax, ty * kBigTileH + kBigTileH - 1);
				mark_rect(p, marking, x0, x1, y0, y1, lane);
				raster_coop(t, p.vis, p.W, lane, x0, x1, y0, y1);
			}
		}
#else
		for (;;) {
			uint32_t w = 0;
			if (lane == 0) w = atomicAdd(p.bigNext, 1u);
			w = __shfl_sync(0xffffffffu, w, 0);
			if (w >= nTiles) break;
			// last record with tileBase <= w: a 32-way search, every lane probing one position per round (4 dependent L2 round trips
			// for a million records instead of 20 — the drain of a scene full of medium triangles is bound by this latency)
			uint32_t lo = 0, hi = nRec - 1;
			while (lo < hi) {
				const uint32_t span = hi - lo, step = span / 32u + 1u;          // probes lo + step, lo + 2*step, ... (clamped to hi)
				const uint32_t probe = min(hi, lo + (lane + 1u) * step);
				const uint32_t ok = __ballot_sync(0xffffffffu, p.big[probe].tileBase <= w); // monotone: a prefix of the lanes
				const uint32_t k = __popc(ok);                                   // probes 1..k hold, probe k+1 (if any) does not
				const uint32_t newLo = k ? min(hi, lo + k * step) : lo;
				const uint32_t newHi = k < 32u ? min(hi, lo + (k + 1u) * step) - 1u : hi;
				lo = newLo; hi = newHi < newLo ? newLo : newHi;
			}
			const BigTri* b = p.big + lo;
			__syncwarp();
			if (lane < sizeof(Tri) / 4) ((uint32_t*)&sTri[warp])[lane] = ((const uint32_t*)&b->t)[lane];
			const uint32_t local = w - b->tileBase, tilesX = b->tilesX;
			__syncwarp();
			const Tri& t = sTri[warp];
			const int tx = t.xmin / kBigTileW + (int)(local % tilesX), ty = t.ymin / kBigTileH + (int)(local / tilesX);
			const int x0 = max(t.xmin, tx * kBigTileW), x1 = min(t.xmax, tx * kBigTileW + kBigTileW - 1);
			const int y0 = max(t.ymin, ty * kBigTileH), y1 = min(t.ymax, ty * kBigTileH + kBigTileH - 1);
			mark_rect(p, marking, x0, x1, y0, y1, lane);
			raster_coop(t, p.vis, p.W, lane, x0, x1, y0, y1);
		}
#endif
	}

#ifndef VKV_DRAIN_NO_REWALK
	if (overflowed) { // a queue was full: rare, slow, correct
		__shared__ WarpScratch slow;
		__shared__ SlowScratch slowSub;
		if (warp == 0) rewalk(p, slow, &slowSub, lane);
	}
#endif
}

__global__ void fill64_kernel(ulonglong2* __restrict__ dst, size_t n2, unsigned long long v, unsigned long long* tail, size_t ntail) {
	const ulonglong2 vv = make_ulonglong2(v, v);
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = vv;
	if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = v;
}
__global__ void fill32_kernel(uint32_t* __restrict__ dst, size_t n, uint32_t v) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void split_vis_kernel(const unsigned long long* __restrict__ vis, size_t n, uint32_t* __restrict__ ids, float* __restrict__ depth) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned long long k = vis[i];
		if (ids) ids[i] = (uint32_t)k;
		if (depth) depth[i] = depth_of_key(k);
	}
}

} // namespace

cudaError_t launch_prepare_transforms(const float* transforms, const vkv_Camera* camera, uint32_t n, float* mvp, uint32_t* detNeg, float4* eye,
                                      int num_sms, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	uint32_t grid = (n + 127) / 128;
	if (grid > (uint32_t)num_sms * 8) grid = (uint32_t)num_sms * 8;
	prepare_transforms_kernel<<<grid, 128, 0, stream>>>(transforms, camera, n, mvp, detNeg, eye);
	return cudaGetLastError();
}

cudaError_t launch_raster(const RasterParams& p, int num_sms, cudaStream_t stream, bool after_cull, bool small_drain) {
	static int perSmOf[64][2] = {}; // per device (a process may hold contexts on several GPUs): hot kernel, drain kernel
	int dev = 0;
	cudaGetDevice(&dev);
	int* perSm = perSmOf[dev & 63];
	if (perSm[0] == 0) {
		int n = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raster_kernel, kThreads, 0);
		perSm[0] = n < 1 ? 1 : n;
		n = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raster_big_kernel, kDrainThreads, 0);
		perSm[1] = n < 1 ? 1 : (n > 4 ? 4 : n);
		if (const char* e = getenv("VKV_DRAIN_BLOCKS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < perSm[1]) perSm[1] = v; } // measurement switch
	}
	cudaLaunchAttribute pdl[1];
	pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	pdl[0].val.programmaticStreamSerializationAllowed = 1;
	if (after_cull) {
		cudaLaunchConfig_t hot = {};
		hot.gridDim = dim3(num_sms * perSm[0]); hot.blockDim = dim3(kThreads); hot.dynamicSmemBytes = 0; hot.stream = stream;
		hot.attrs = pdl; hot.numAttrs = 1;
		cudaError_t e = cudaLaunchKernelEx(&hot, raster_kernel, p);
		if (e != cudaSuccess) return e;
	} else raster_kernel<<<num_sms * perSm[0], kThreads, 0, stream>>>(p);
	// every block co-resident (its clip phase ends in a grid barrier); exits at once when both queues are empty.  Launched with
	// programmatic stream serialization: its blocks are set up under raster_kernel's tail (see there) instead of after it.
	cudaLaunchConfig_t cfg = {};
	// small_drain: the caller has seen this pass's queues empty in the frames before.  An empty drain is pure latency between the rasteriser and
	// the pyramid build, and most of it is block turnover: 444 / 296 / 148 / 74 blocks cost 15 / 12 / 7 / 5 us per frame on cfg 3 (profiles/r4m).
	// Half a block per SM still drains an occasional clipped or large triangle; a scene that fills the queues gets the full grid back from
	// the next observed frame on (cfg 2 with one block per SM: 0.260 against 0.217 ms).
	cfg.gridDim = dim3(small_drain ? max(1, num_sms / 2) : num_sms * perSm[1]); cfg.blockDim = dim3(kDrainThreads); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
	static const int gridOverride = getenv("VKV_DRAIN_GRID") ? atoi(getenv("VKV_DRAIN_GRID")) : 0; // measurement switch
	if (gridOverride > 0 && gridOverride < (int)cfg.gridDim.x) cfg.gridDim = dim3(gridOverride);
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, raster_big_kernel, p);
}

cudaError_t launch_fill64(unsigned long long* dst, size_t n, unsigned long long value, int num_sms, cudaStream_t stream) {
	const size_t n2 = n / 2;
	size_t grid = (n2 + 255) / 256;
	if (grid > (size_t)num_sms * 16) grid = (size_t)num_sms * 16;
	if (grid == 0) grid = 1;
	fill64_kernel<<<(unsigned)grid, 256, 0, stream>>>((ulonglong2*)dst, n2, value, dst + n2 * 2, n - n2 * 2);
	return cudaGetLastError();
}
cudaError_t launch_fill32(uint32_t* dst, size_t n, uint32_t value, int num_sms, cudaStream_t stream) {
	size_t grid = (n + 255) / 256;
	if (grid > (size_t)num_sms * 16) grid = (size_t)num_sms * 16;
	if (grid == 0) grid = 1;
	fill32_kernel<<<(unsigned)grid, 256, 0, stream>>>(dst, n, value);
	return cudaGetLastError();
}
cudaError_t launch_split_vis(const unsigned long long* vis, size_t n, uint32_t* ids, float* depth, int num_sms, cudaStream_t stream) {
	size_t grid = (n + 255) / 256;
	if (grid > (size_t)num_sms * 16) grid = (size_t)num_sms * 16;
	if (grid == 0) grid = 1;
	split_vis_kernel<<<(unsigned)grid, 256, 0, stream>>>(vis, n, ids, depth);
	return cudaGetLastError();
}
