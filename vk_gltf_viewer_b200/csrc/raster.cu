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
constexpr int kBatch = 8;          // meshlets fetched per work-stealing grab (one lane walks each meshlet's header chain)
constexpr int kMinBlocks = 3;      // register budget: 65536 / (256 * 3) = 85 -> 24 warps / SM

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

template <bool PRE_READ>
__device__ __forceinline__ void shade(unsigned long long* __restrict__ vis, uint32_t W, int x, int y, float w1, float w2, float invA,
                                      float za, float dzb, float dzc, uint32_t id) {
	const float l1 = w1 * invA;
	const float l2 = w2 * invA;
	float z = (za + l1 * dzb) + l2 * dzc;
	z = (z > 0.0f) ? z : 0.0f;
	z = (z < 1.0f) ? z : 1.0f;
	const unsigned long long key = ((unsigned long long)(~__float_as_uint(z)) << 32) | id;
	unsigned long long* p = vis + (size_t)y * W + x;
	// PRE_READ: a plain L2 read filters fragments that are already behind (saves atomic traffic, costs an L2 round trip
	// per fragment); otherwise a fire-and-forget RED.MIN.U64 — the issuing lane never waits.
	if (PRE_READ) { if (key < __ldcg(p)) atomicMin(p, key); }
	else atomicMin(p, key);
}

// Lane-serial scan of a small triangle; all edge values fit in int32 (deltas <= 2^14 sub-pixels).
template <bool PRE_READ>
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
			if (w0 >= b0 && w1 >= b1 && w2 >= b2) shade<PRE_READ>(vis, W, x, y, (float)w1, (float)w2, t.invA, t.za, t.dzb, t.dzc, t.id);
			w0 -= e0dy * VKV_SUB; w1 -= e1dy * VKV_SUB; w2 -= e2dy * VKV_SUB;
		}
		r0 += e0dx * VKV_SUB; r1 += e1dx * VKV_SUB; r2 += e2dx * VKV_SUB;
	}
}

// Whole-warp scan in 8x4 stamps with stamp-level rejection; int64 edge functions (any triangle inside the guard band).
template <bool PRE_READ>
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
					shade<PRE_READ>(vis, W, x, y, (float)w1, (float)w2, t.invA, t.za, t.dzb, t.dzc, t.id);
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

// Per-transform prologue of the mesh shader (mesh.glsl:43-44,71), hoisted out of the per-meshlet path:
// mvp = viewProjection * transform (column by column) and the sign of determinant(transform).  Same arithmetic as before,
// computed once per mesh-node instead of once per meshlet.
__global__ void prepare_transforms_kernel(const float* __restrict__ transforms, const vkv_Camera* __restrict__ camera, uint32_t n,
                                          float* __restrict__ mvpOut, uint32_t* __restrict__ detNeg) {
	__shared__ float sVP[16];
	if (threadIdx.x < 16) sVP[threadIdx.x] = __ldg(camera->viewProjection + threadIdx.x);
	__syncthreads();
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
		float tm[16];
#pragma unroll
		for (int c = 0; c < 4; ++c) {
			const float4 col = __ldg((const float4*)(transforms + (size_t)t * 16 + c * 4));
			tm[c * 4] = col.x; tm[c * 4 + 1] = col.y; tm[c * 4 + 2] = col.z; tm[c * 4 + 3] = col.w;
		}
#pragma unroll
		for (int c = 0; c < 4; ++c) {
			const float4 r = mul44(sVP, tm[c * 4], tm[c * 4 + 1], tm[c * 4 + 2], tm[c * 4 + 3]);
			*(float4*)(mvpOut + (size_t)t * 16 + c * 4) = r;
		}
		detNeg[t] = det4(tm) < 0.0f ? 1u : 0u;
	}
}

// what one lane fetches for one meshlet of a batch (mesh.glsl:31-36 resolved to addresses)
struct alignas(16) MeshletHdr {
	const uint32_t* vidx;      // primitive.vertexIndexBuffer + meshlet.vertexOffset
	const uint8_t* tri;        // primitive.primitiveIndexBuffer + meshlet.triangleOffset
	const vkv_Vertex* verts;   // primitive.vertexBuffer
	uint32_t drawId;
	uint32_t tIdx;
	uint32_t counts;           // vertexCount | triangleCount << 8 | doubleSided << 16 | detNegative << 17
	uint32_t pad[3];
};

template <bool PRE_READ>
__global__ void __launch_bounds__(kThreads, kMinBlocks) raster_kernel(const RasterParams p) {
	__shared__ WarpScratch scratch[kWarpsPerBlock];
	__shared__ MeshletHdr hdrs[kWarpsPerBlock][kBatch];
	__shared__ __align__(16) float sMvp[kWarpsPerBlock][16];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	WarpScratch& ws = scratch[warp];
	const uint32_t count = __ldg(p.count);
	const float hw = (float)p.W * 0.5f, hh = (float)p.H * 0.5f;

	for (;;) {
		uint32_t base = 0;
		if (lane == 0) base = atomicAdd(p.work, (uint32_t)kBatch);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= count) break;
		const uint32_t nb = min((uint32_t)kBatch, count - base);
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
			h.verts = (const vkv_Vertex*)b1.x;
			h.drawId = drawId; h.tIdx = tIdx;
			h.counts = vc | (tc << 8) | (ds << 16) | (dn << 17);
			hdrs[warp][lane] = h;
		}
		__syncwarp();

		// software pipeline over the batch: while meshlet j is processed, the vertex indices (then positions), the triangle
		// bytes and the mvp of meshlet j+1 are already in flight
		uint32_t vi0 = 0, vi1 = 0;                 // vertex indices of the NEXT meshlet (slots lane, lane+32)
		float px0 = 0, py0 = 0, pz0 = 0, px1 = 0, py1 = 0, pz1 = 0; // positions of the CURRENT meshlet
		uint32_t tw0 = 0, tw1 = 0, tw2 = 0;        // triangle index words of the CURRENT meshlet
		float mv = 0.f;                            // lanes 0..15: mvp element of the CURRENT meshlet
		{
			const MeshletHdr& h = hdrs[warp][0];
			const uint32_t vc = h.counts & 0xffu;
			if (lane < vc) vi0 = __ldg(h.vidx + lane);
			if (lane + 32 < vc) vi1 = __ldg(h.vidx + lane + 32);
		}
		auto issue_loads = [&](const MeshletHdr& h) { // positions + triangle words + mvp of meshlet h (vi0/vi1 hold its indices)
			const uint32_t vc = h.counts & 0xffu, tc = (h.counts >> 8) & 0xffu;
			if (lane < vc) { const float* q = h.verts[vi0].position; px0 = __ldg(q); py0 = __ldg(q + 1); pz0 = __ldg(q + 2); }
			if (lane + 32 < vc) { const float* q = h.verts[vi1].position; px1 = __ldg(q); py1 = __ldg(q + 1); pz1 = __ldg(q + 2); }
			const uint32_t nWords = (tc * 3 + 3) >> 2;
			if ((((uintptr_t)h.tri) & 3) == 0) {
				const uint32_t* w = (const uint32_t*)h.tri;
				if (lane < nWords) tw0 = __ldg(w + lane);
				if (lane + 32 < nWords) tw1 = __ldg(w + lane + 32);
				if (lane + 64 < nWords) tw2 = __ldg(w + lane + 64);
			} else { // unaligned triangle slice (never produced by the reference's builder: assets.cpp:339 pads to 4)
				const uint32_t nBytes = tc * 3;
				auto gather = [&](uint32_t wi) {
					uint32_t r = 0;
					for (uint32_t b = 0; b < 4; ++b) if (wi * 4 + b < nBytes) r |= (uint32_t)__ldg(h.tri + wi * 4 + b) << (8 * b);
					return r;
				};
				if (lane < nWords) tw0 = gather(lane);
				if (lane + 32 < nWords) tw1 = gather(lane + 32);
				if (lane + 64 < nWords) tw2 = gather(lane + 64);
			}
			if (lane < 16) mv = __ldg(p.mvp + (size_t)h.tIdx * 16 + lane);
		};
		issue_loads(hdrs[warp][0]);

		for (uint32_t j = 0; j < nb; ++j) {
			const MeshletHdr h = hdrs[warp][j];
			const uint32_t drawId = h.drawId;
			const uint32_t vc = h.counts & 0xffu, tc = (h.counts >> 8) & 0xffu;
			const bool doubleSided = (h.counts >> 16) & 1u, detNeg = (h.counts >> 17) & 1u;
			// vertex indices of meshlet j+1: in flight during this meshlet's vertex phase
			if (j + 1 < nb) {
				const MeshletHdr& hn = hdrs[warp][j + 1];
				const uint32_t vcn = hn.counts & 0xffu;
				if (lane < vcn) vi0 = __ldg(hn.vidx + lane);
				if (lane + 32 < vcn) vi1 = __ldg(hn.vidx + lane + 32);
			}
			// stage this meshlet's triangle words and mvp
			ws.tri_words[lane] = tw0; ws.tri_words[lane + 32] = tw1; ws.tri_words[lane + 64] = tw2;
			if (lane < 16) sMvp[warp][lane] = mv;
			__syncwarp();
			float mvp[16];
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const float4 col = *(const float4*)&sMvp[warp][c * 4];
				mvp[c * 4] = col.x; mvp[c * 4 + 1] = col.y; mvp[c * 4 + 2] = col.z; mvp[c * 4 + 3] = col.w;
			}
			// :50-69 vertices
#pragma unroll
			for (int half = 0; half < 2; ++half) {
				const uint32_t v = lane + half * 32;
				if (v < vc) {
					const float4 c = half ? mul44(mvp, px1, py1, pz1, 1.0f) : mul44(mvp, px0, py0, pz0, 1.0f); // :61
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
			}
			__syncwarp();
			// positions / triangle words / mvp of meshlet j+1: in flight during this meshlet's triangle phase
			if (j + 1 < nb) issue_loads(hdrs[warp][j + 1]);

			// :73-103 triangles
			const uint8_t* tb = (const uint8_t*)ws.tri_words;
			for (uint32_t tbase = 0; tbase < tc; tbase += 32) {
				const uint32_t t = tbase + lane;
				int kind = 0; // 0 nothing, 1 serial, 2 cooperative, 3 clip
				Tri tri;
				uint32_t ia = 0, ib = 0, ic = 0;
				if (t < tc) {
					ia = tb[t * 3]; ib = tb[t * 3 + 1]; ic = tb[t * 3 + 2];
					ia = min(ia, vc - 1); ib = min(ib, vc - 1); ic = min(ic, vc - 1); // robustness only
					const float4 A = ws.clip[ia], B = ws.clip[ib], C = ws.clip[ic];
					const uint32_t fa = ws.flags[ia], fb = ws.flags[ib], fc = ws.flags[ic];
					bool cull = false;
					if (!doubleSided) { // :86-98
						const float det = det3(make_float3(A.x, A.y, A.w), make_float3(B.x, B.y, B.w), make_float3(C.x, C.y, C.w));
						cull = detNeg ? (det < 0.0f) : (det > 0.0f);
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
				if (kind == 1) raster_serial<PRE_READ>(tri, p.vis, p.W);
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
					for (int s = 0; s < n; ++s) raster_coop<PRE_READ>(ws.sub[s], p.vis, p.W, lane);
					__syncwarp();
				}
			}
			__syncwarp();
		}
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

cudaError_t launch_prepare_transforms(const float* transforms, const vkv_Camera* camera, uint32_t n, float* mvp, uint32_t* detNeg,
                                      int num_sms, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	uint32_t grid = (n + 127) / 128;
	if (grid > (uint32_t)num_sms * 8) grid = (uint32_t)num_sms * 8;
	prepare_transforms_kernel<<<grid, 128, 0, stream>>>(transforms, camera, n, mvp, detNeg);
	return cudaGetLastError();
}

cudaError_t launch_raster(const RasterParams& p, int num_sms, cudaStream_t stream) {
	static int perSm[2] = {0, 0};
	const int v = p.pre_read ? 1 : 0;
	if (perSm[v] == 0) {
		int n = 0;
		if (v) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raster_kernel<true>, kThreads, 0);
		else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raster_kernel<false>, kThreads, 0);
		perSm[v] = n < 1 ? 1 : n;
	}
	if (v) raster_kernel<true><<<num_sms * perSm[v], kThreads, 0, stream>>>(p);
	else raster_kernel<false><<<num_sms * perSm[v], kThreads, 0, stream>>>(p);
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
