// raster.cu — software visibility-buffer rasteriser (sm_100a).
// Replaces, for every surviving MeshletDraw: shaders/visbuffer/visbuffer.mesh.glsl:30-104 (vertex transform +
// per-triangle facing cull), the fixed-function clip / viewport / raster / depth stages configured at
// application.cpp:326-340,772-841 (+ src/vulkan/pipeline_builder.cpp:225-277) and visbuffer.frag.glsl:36.
//
// One warp per meshlet (work-stealing over the survivor list).  Vertices are transformed once into per-warp shared
// memory (clip position + snapped screen position); each lane then owns triangles: facing cull, trivial reject,
// integer edge setup.  Small triangles are scanned by their lane (int32 edge functions — exact, no overflow);
// large or clipped ones are handed to the whole warp (8x4 pixel stamps, int64 edge functions, stamp-level reject).
// Visibility is resolved with ONE 64-bit atomicMin per covered pixel on (~depthBits << 32 | drawId << 7 | tri),
// preceded by a plain read that filters already-occluded fragments.
#include "kernels.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kSerialMaxDim = 8;   // lane-serial path: bbox <= 8x8 pixels

struct Tri {
	int ax, ay, bx, by, cx, cy;      // snapped vertices, 24.8 fixed point, area2 > 0
	int xmin, xmax, ymin, ymax;      // pixel bbox clipped to the viewport
	long long area2;
	float za, dzb, dzc, invA;
	uint32_t id;
	uint32_t small;                  // vertex extent <= 2^14 sub-pixels in x and y: every edge value fits in int32
};

struct WarpScratch {
	float4 clip[VKV_MAX_VERTICES];
	int2 fxy[VKV_MAX_VERTICES];
	float zndc[VKV_MAX_VERTICES];
	uint32_t flags[VKV_MAX_VERTICES];
	uint32_t tri_words[96];          // up to 124*3 = 372 index bytes
	Tri sub[8];
	int nsub;
};

enum { F_NEEDS_CLIP = 64 };

__device__ __forceinline__ bool top_left(int dx, int dy) { return dy < 0 || (dy == 0 && dx > 0); }

__device__ __forceinline__ void project(float4 c, float hw, float hh, int& fx, int& fy, float& z) {
	const float nx = c.x / c.w, ny = c.y / c.w;
	z = c.z / c.w;
	const float sx = nx * hw + hw;
	const float sy = ny * hh + hh;
	fx = __float2int_rn(sx * (float)VKV_SUB);
	fy = __float2int_rn(sy * (float)VKV_SUB);
}

// integer setup shared by every path; false = nothing to draw
__device__ __forceinline__ bool setup_tri(int ax, int ay, float za, int bx, int by, float zb, int cx, int cy, float zc, uint32_t id,
                                          uint32_t W, uint32_t H, Tri& t) {
	long long area2 = (long long)(bx - ax) * (cy - ay) - (long long)(by - ay) * (cx - ax);
	if (area2 == 0) return false;
	if (area2 < 0) { // cullMode NONE: both windings are drawn
		int tx = bx; bx = cx; cx = tx;
		int ty = by; by = cy; cy = ty;
		float tz = zb; zb = zc; zc = tz;
		area2 = -area2;
	}
	const int minx = min(ax, min(bx, cx)), maxx = max(ax, max(bx, cx));
	const int miny = min(ay, min(by, cy)), maxy = max(ay, max(by, cy));
	t.xmin = max(0, (minx + (VKV_SUB / 2 - 1)) >> VKV_SUB_BITS);
	t.xmax = min((int)W - 1, (maxx - VKV_SUB / 2) >> VKV_SUB_BITS);
	t.ymin = max(0, (miny + (VKV_SUB / 2 - 1)) >> VKV_SUB_BITS);
	t.ymax = min((int)H - 1, (maxy - VKV_SUB / 2) >> VKV_SUB_BITS);
	if (t.xmin > t.xmax || t.ymin > t.ymax) return false;
	t.ax = ax; t.ay = ay; t.bx = bx; t.by = by; t.cx = cx; t.cy = cy;
	t.area2 = area2;
	t.za = za; t.dzb = zb - za; t.dzc = zc - za;
	t.invA = 1.0f / (float)area2;
	t.id = id;
	t.small = (maxx - minx <= 16384 && maxy - miny <= 16384) ? 1u : 0u;
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
	unsigned long long* p = vis + (size_t)y * W + x;
	if (key < __ldcg(p)) atomicMin(p, key);
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

// Whole-warp scan in 8x4 stamps with stamp-level rejection; int64 edge functions (any triangle inside the guard band).
__device__ __noinline__ void raster_coop(const Tri& t, unsigned long long* __restrict__ vis, uint32_t W, uint32_t lane) {
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
	const long long px0 = (long long)t.xmin * VKV_SUB + VKV_SUB / 2, py0 = (long long)t.ymin * VKV_SUB + VKV_SUB / 2;
	long long r0 = e0dx * (py0 - t.by) - e0dy * (px0 - t.bx);
	long long r1 = e1dx * (py0 - t.cy) - e1dy * (px0 - t.cx);
	long long r2 = e2dx * (py0 - t.ay) - e2dy * (px0 - t.ax);
	for (int ty = t.ymin; ty <= t.ymax; ty += 4) {
		long long s0 = r0, s1 = r1, s2 = r2;
		for (int tx = t.xmin; tx <= t.xmax; tx += 8) {
			if (s0 + m0 >= b0 && s1 + m1 >= b1 && s2 + m2 >= b2) {
				const long long w0 = s0 + o0, w1 = s1 + o1, w2 = s2 + o2;
				const int x = tx + lx, y = ty + ly;
				if (x <= t.xmax && y <= t.ymax && w0 >= b0 && w1 >= b1 && w2 >= b2)
					shade(vis, W, x, y, (float)w1, (float)w2, t.invA, t.za, t.dzb, t.dzc, t.id);
			}
			s0 -= e0dy * 8 * VKV_SUB; s1 -= e1dy * 8 * VKV_SUB; s2 -= e2dy * 8 * VKV_SUB;
		}
		r0 += e0dx * 4 * VKV_SUB; r1 += e1dx * 4 * VKV_SUB; r2 += e2dx * 4 * VKV_SUB;
	}
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

__global__ void __launch_bounds__(kThreads) raster_kernel(const RasterParams p) {
	__shared__ WarpScratch scratch[kWarpsPerBlock];
	__shared__ float sVP[16];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	WarpScratch& ws = scratch[warp];
	if (threadIdx.x < 16) sVP[threadIdx.x] = __ldg(p.camera->viewProjection + threadIdx.x);
	__syncthreads();
	const uint32_t count = __ldg(p.count);
	const float hw = (float)p.W * 0.5f, hh = (float)p.H * 0.5f;

	for (;;) {
		uint32_t item = 0;
		if (lane == 0) item = atomicAdd(p.work, 1u);
		item = __shfl_sync(0xffffffffu, item, 0);
		if (item >= count) break;
		const uint32_t drawId = __ldg(p.list + item);
		// mesh.glsl:31-36
		const vkv_MeshletDraw* d = p.draws + drawId;
		const uint32_t primIdx = __ldg(&d->primitiveIndex), mlIdx = __ldg(&d->meshletIndex), tIdx = __ldg(&d->transformIndex);
		const vkv_Primitive* prim = p.primitives + primIdx;
		const vkv_Meshlet* ml = (const vkv_Meshlet*)__ldg(&prim->meshletBuffer) + mlIdx;
		const uint32_t vertexOffset = __ldg(&ml->vertexOffset), triangleOffset = __ldg(&ml->triangleOffset);
		const uint32_t counts = __ldg((const uint32_t*)&ml->vertexCount);
		const uint32_t vc = min(counts & 0xffu, VKV_MAX_VERTICES), tc = min((counts >> 8) & 0xffu, VKV_MAX_MESHLET_TRIANGLES);
		const bool doubleSided = __ldg(&p.materials[__ldg(&prim->materialIndex)].doubleSided) != 0;
		const float* T = p.transforms + (size_t)tIdx * 16;
		// :43-44 mvp = viewProjection * transform, column by column
		float tm[16], mvp[16];
#pragma unroll
		for (int c = 0; c < 4; ++c) {
			const float4 col = __ldg((const float4*)(T + c * 4));
			tm[c * 4] = col.x; tm[c * 4 + 1] = col.y; tm[c * 4 + 2] = col.z; tm[c * 4 + 3] = col.w;
		}
#pragma unroll
		for (int c = 0; c < 4; ++c) {
			const float4 r = mul44(sVP, tm[c * 4], tm[c * 4 + 1], tm[c * 4 + 2], tm[c * 4 + 3]);
			mvp[c * 4] = r.x; mvp[c * 4 + 1] = r.y; mvp[c * 4 + 2] = r.z; mvp[c * 4 + 3] = r.w;
		}
		const float transformDet = det4(tm); // :71

		// stage triangle index bytes (coalesced words when the slice is 4-byte aligned)
		const uint8_t* triBytes = (const uint8_t*)__ldg(&prim->primitiveIndexBuffer) + triangleOffset;
		const uint32_t nTriBytes = tc * 3;
		if ((((uintptr_t)triBytes) & 3) == 0) {
			for (uint32_t i = lane; i < (nTriBytes + 3) / 4; i += 32) ws.tri_words[i] = __ldg((const uint32_t*)triBytes + i);
		} else {
			uint8_t* dst = (uint8_t*)ws.tri_words;
			for (uint32_t i = lane; i < nTriBytes; i += 32) dst[i] = __ldg(triBytes + i);
		}
		// :50-69 vertices
		const uint32_t* vidx = (const uint32_t*)__ldg(&prim->vertexIndexBuffer) + vertexOffset;
		const vkv_Vertex* verts = (const vkv_Vertex*)__ldg(&prim->vertexBuffer);
		for (uint32_t v = lane; v < vc; v += 32) {
			const float* pos = verts[__ldg(vidx + v)].position;
			const float4 c = mul44(mvp, __ldg(pos), __ldg(pos + 1), __ldg(pos + 2), 1.0f); // :61
			uint32_t f = 0;
			if (c.x < -c.w) f |= 1;
			if (c.x > c.w) f |= 2;
			if (c.y < -c.w) f |= 4;
			if (c.y > c.w) f |= 8;
			if (c.z < 0.f) f |= 16;
			if (c.z > c.w) f |= 32;
			const float g = VKV_GUARD * c.w;
			if (c.z < 0.f || c.z > c.w || c.x > g || c.x < -g || c.y > g || c.y < -g) f |= F_NEEDS_CLIP;
			if (!(c.x == c.x && c.y == c.y && c.z == c.z && c.w == c.w)) f |= 0x80; // NaN -> reject
			int fx = 0, fy = 0;
			float z = 0.f;
			if (!(f & (F_NEEDS_CLIP | 0x80)) && c.w > 0.f) project(c, hw, hh, fx, fy, z);
			else if (!(f & 0x80) && !(c.w > 0.f)) f |= F_NEEDS_CLIP; // degenerate w: let the clipper decide
			ws.clip[v] = c;
			ws.fxy[v] = make_int2(fx, fy);
			ws.zndc[v] = z;
			ws.flags[v] = f;
		}
		__syncwarp();

		// :73-103 triangles
		const uint8_t* tb = (const uint8_t*)ws.tri_words;
		for (uint32_t base = 0; base < tc; base += 32) {
			const uint32_t t = base + lane;
			int kind = 0; // 0 nothing, 1 serial, 2 cooperative, 3 clip
			Tri tri;
			uint32_t ia = 0, ib = 0, ic = 0;
			if (t < tc) {
				ia = tb[t * 3]; ib = tb[t * 3 + 1]; ic = tb[t * 3 + 2];
				if (ia >= vc) ia = vc - 1; if (ib >= vc) ib = vc - 1; if (ic >= vc) ic = vc - 1; // robustness only
				const float4 A = ws.clip[ia], B = ws.clip[ib], C = ws.clip[ic];
				const uint32_t fa = ws.flags[ia], fb = ws.flags[ib], fc = ws.flags[ic];
				bool cull = false;
				if (!doubleSided) { // :86-98
					const float det = det3(make_float3(A.x, A.y, A.w), make_float3(B.x, B.y, B.w), make_float3(C.x, C.y, C.w));
					cull = (transformDet < 0.0f) ? (det < 0.0f) : (det > 0.0f);
				}
				const uint32_t id = (drawId << VKV_TRIANGLE_BITS) | t; // frag.glsl:36
				if (!cull && !((fa | fb | fc) & 0x80) && !(fa & fb & fc & 63)) {
					if ((fa | fb | fc) & F_NEEDS_CLIP) kind = 3;
					else {
						const int2 a = ws.fxy[ia], b = ws.fxy[ib], c = ws.fxy[ic];
						if (setup_tri(a.x, a.y, ws.zndc[ia], b.x, b.y, ws.zndc[ib], c.x, c.y, ws.zndc[ic], id, p.W, p.H, tri))
							kind = (tri.small && tri.xmax - tri.xmin < kSerialMaxDim && tri.ymax - tri.ymin < kSerialMaxDim) ? 1 : 2;
					}
				}
			}
			if (kind == 1) raster_serial(tri, p.vis, p.W);
			uint32_t coop = __ballot_sync(0xffffffffu, kind >= 2);
			while (coop) {
				const int src = __ffs(coop) - 1;
				coop &= coop - 1;
				if ((int)lane == src) {
					if (kind == 2) { ws.sub[0] = tri; ws.nsub = 1; }
					else ws.nsub = clip_and_setup(ws.clip[ia], ws.clip[ib], ws.clip[ic], (drawId << VKV_TRIANGLE_BITS) | t, p.W, p.H, ws.sub);
				}
				__syncwarp();
				const int n = ws.nsub;
				for (int s = 0; s < n; ++s) raster_coop(ws.sub[s], p.vis, p.W, lane);
				__syncwarp();
			}
		}
		__syncwarp();
	}
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

cudaError_t launch_raster(const RasterParams& p, int num_sms, cudaStream_t stream) {
	int perSm = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, raster_kernel, kThreads, 0);
	if (perSm < 1) perSm = 1;
	raster_kernel<<<num_sms * perSm, kThreads, 0, stream>>>(p);
	return cudaGetLastError();
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
