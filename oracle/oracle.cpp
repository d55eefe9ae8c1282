// oracle.cpp — CPU restatement of the reference's cull / visbuffer raster / HiZ path.
// TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED except where oracle.h says otherwise.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
//
// Every function cites the reference file:line it restates (paths relative to the upstream tree).
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <limits>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Arithmetic policy.  Every operation is a single correctly-rounded fp32 op.
// ---------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline float gmin(float x, float y) { return y < x ? y : x; } // GLSL min(x,y)
inline float gmax(float x, float y) { return x < y ? y : x; } // GLSL max(x,y)
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }

// mat4 (column-major m[c*4+r]) * vec4 : ((c0*x + c1*y) + c2*z) + c3*w
inline V4 mul44(const float* m, V4 v) {
	V4 r;
	r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
	r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
	r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
	r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
	return r;
}
inline void mul44m(const float* a, const float* b, float* out) { // out = a*b, column by column
	for (int c = 0; c < 4; ++c) {
		V4 col = mul44(a, V4{b[c * 4 + 0], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]});
		out[c * 4 + 0] = col.x; out[c * 4 + 1] = col.y; out[c * 4 + 2] = col.z; out[c * 4 + 3] = col.w;
	}
}
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// determinant(mat3(c0,c1,c2)) = dot(c0, c1.yzx*c2.zxy - c1.zxy*c2.yzx)
inline float det3(V3 c0, V3 c1, V3 c2) {
	V3 d;
	d.x = c1.y * c2.z - c1.z * c2.y;
	d.y = c1.z * c2.x - c1.x * c2.z;
	d.z = c1.x * c2.y - c1.y * c2.x;
	return dot3(c0, d);
}
// determinant(mat4): cofactor expansion along the first column; only the sign is consumed (mesh.glsl:71,94)
inline float det4(const float* m) {
	auto col3 = [&](int c, int skipRow) {
		float v[3]; int k = 0;
		for (int r = 0; r < 4; ++r) if (r != skipRow) v[k++] = m[c * 4 + r];
		return V3{v[0], v[1], v[2]};
	};
	float d0 = det3(col3(1, 0), col3(2, 0), col3(3, 0));
	float d1 = det3(col3(1, 1), col3(2, 1), col3(3, 1));
	float d2 = det3(col3(1, 2), col3(2, 2), col3(3, 2));
	float d3 = det3(col3(1, 3), col3(2, 3), col3(3, 3));
	return ((m[0] * d0 - m[1] * d1) + m[2] * d2) - m[3] * d3;
}

inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

inline bool near_rel(float a, float b, float ulps) {
	float m = std::fmax(std::fabs(a), std::fabs(b));
	return std::fabs(a - b) <= ulps * 1.1920929e-7f * m;
}

template <class F>
void parallel_for(size_t n, int threads, F fn) {
	if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
	if (threads < 1) threads = 1;
	if ((size_t)threads > n) threads = n ? (int)n : 1;
	if (threads == 1) { fn(0, n, 0); return; }
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; ++t) {
		size_t b = n * t / threads, e = n * (t + 1) / threads;
		pool.emplace_back([=] { fn(b, e, t); });
	}
	for (auto& th : pool) th.join();
}

// ---------------------------------------------------------------------------------------------
// Min-reduction LINEAR sampler with CLAMP_TO_EDGE (application.cpp:438-453):
// texels {i0, i0+1} per axis with i0 = floor(u*size - 0.5); a texel whose bilinear weight is 0
// (frac == 0 -> i0+1) takes no part in the reduction.
// ---------------------------------------------------------------------------------------------
inline void footprint(float coord, uint32_t size, int& lo, int& hi, int* ambig) {
	float u = coord * (float)size - 0.5f;
	if (!(u >= -1.0f)) { lo = hi = 0; return; }                      // NaN or far left: clamps to texel 0
	if (u >= (float)size) { lo = hi = (int)size - 1; return; }
	float fl = std::floor(u);
	float frac = u - fl;
	int i0 = (int)fl;
	int i1 = (frac == 0.0f) ? i0 : i0 + 1;
	if (ambig && (frac < 1e-4f || frac > 1.0f - 1e-4f)) *ambig |= 1;
	int mx = (int)size - 1;
	lo = i0 < 0 ? 0 : (i0 > mx ? mx : i0);
	hi = i1 < 0 ? 0 : (i1 > mx ? mx : i1);
}

inline float sample_min(const float* img, uint32_t w, uint32_t h, float u, float v, int* ambig) {
	int x0, x1, y0, y1;
	footprint(u, w, x0, x1, ambig);
	footprint(v, h, y0, y1, ambig);
	float m = img[(size_t)y0 * w + x0];
	m = gmin(m, img[(size_t)y0 * w + x1]);
	m = gmin(m, img[(size_t)y1 * w + x0]);
	m = gmin(m, img[(size_t)y1 * w + x1]);
	return m;
}

// culling.h.glsl:32-41
const float kAabbPositions[8][3] = {
	{1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}, {1, -1, 1}, {1, 1, 1}, {-1, -1, 1}, {-1, 1, 1}};

struct Pyr {
	uint32_t levels, off[17], w[16], h[16], total;
};

// Diagnostics (the ORC_AMBIG_* flags) cost about as much again as the decision itself; bench.py's timed CPU baseline switches
// them off (orc_set_diagnostics(0)) so that the baseline measures the path, not the bookkeeping.
bool g_diagnostics = true;

// How far a DIFFERENT but equally valid evaluation of the same GLSL (another association order of the mat4*vec4 sums, an
// FMA-contracting compiler) can move a result: kNoise relative roundings of the LARGEST term of each sum, propagated to first
// order.  A result closer than that to one of the shader's thresholds is "ambiguous": flagged, never silently decided.
constexpr float kNoise = 8.0f * 5.9604645e-8f;   // 8 roundings of 2^-24

// visbuffer.task.glsl:44-65 for one MeshletDraw
// Optional normal-cone stage (extension, NOT in the reference — assets.cpp:323 disables cones): the camera position in each
// mesh-node's own space, from rows x, y, w of viewProjection * transform (the same expression, in the same association, as
// common.cuh transform_prologue), and the per-primitive cone arrays.
struct ConeCtx { const uint64_t* table; const V4* eye; };

V4 mesh_space_eye(const float* VP, const float* T) {
	float m[16];
	mul44m(VP, T, m);
	const V3 r0{m[0], m[4], m[8]}, r1{m[1], m[5], m[9]}, r2{m[3], m[7], m[11]};
	const float b0 = -m[12], b1 = -m[13], b2 = -m[15];
	const V3 x12{r1.y * r2.z - r1.z * r2.y, r1.z * r2.x - r1.x * r2.z, r1.x * r2.y - r1.y * r2.x};
	const V3 x20{r2.y * r0.z - r2.z * r0.y, r2.z * r0.x - r2.x * r0.z, r2.x * r0.y - r2.y * r0.x};
	const V3 x01{r0.y * r1.z - r0.z * r1.y, r0.z * r1.x - r0.x * r1.z, r0.x * r1.y - r0.y * r1.x};
	const float D = dot3(r0, x12);
	V4 e;
	e.x = ((b0 * x12.x + b1 * x20.x) + b2 * x01.x) / D;
	e.y = ((b0 * x12.y + b1 * x20.y) + b2 * x01.y) / D;
	e.z = ((b0 * x12.z + b1 * x20.z) + b2 * x01.z) / D;
	e.w = 0.0f;
	if (!(D != 0.0f) || !(std::fabs(D) <= 3.4028234e38f)) e.x = e.y = e.z = std::numeric_limits<float>::quiet_NaN();
	return e;
}
constexpr float kConeMargin = 1.0e-3f; // == common.cuh

uint8_t cull_one(const vkv_VisbufferPushConstants* pc, uint32_t drawIdx, const vkv_Camera& cam, const float* occVP,
                 const Pyr& pyr, const float* pyramid, const ConeCtx* cone = nullptr) {
	const vkv_MeshletDraw* draws = (const vkv_MeshletDraw*)pc->drawBuffer;
	const float* transforms = (const float*)pc->transformBuffer;
	const vkv_Primitive* prims = (const vkv_Primitive*)pc->primitiveBuffer;
	const vkv_MeshletDraw d = draws[drawIdx];                                   // task.glsl:44
	const float* T = transforms + (size_t)d.transformIndex * 16;                // :45
	const vkv_Primitive& prim = prims[d.primitiveIndex];                        // :46
	const vkv_Meshlet& ml = ((const vkv_Meshlet*)prim.meshletBuffer)[d.meshletIndex]; // :47
	uint8_t flags = 0;
	const bool diag = g_diagnostics;

	// :50  worldAabbCenter = (T * vec4(center,1)).xyz
	V4 wc4 = mul44(T, V4{ml.aabbCenter[0], ml.aabbCenter[1], ml.aabbCenter[2], 1.0f});
	V3 wc{wc4.x, wc4.y, wc4.z};
	// :51 -> culling.h.glsl:22-29  abs(mat3(T)) * extent
	V3 e{ml.aabbExtents[0], ml.aabbExtents[1], ml.aabbExtents[2]};
	V3 we;
	we.x = (std::fabs(T[0]) * e.x + std::fabs(T[4]) * e.y) + std::fabs(T[8]) * e.z;
	we.y = (std::fabs(T[1]) * e.x + std::fabs(T[5]) * e.y) + std::fabs(T[9]) * e.z;
	we.z = (std::fabs(T[2]) * e.x + std::fabs(T[6]) * e.y) + std::fabs(T[10]) * e.z;
	// rounding noise of a corner coordinate per axis: the terms of the centre's sum plus the extent
	float posErr[3] = {0.f, 0.f, 0.f};
	if (diag) {
		const float c[3] = {ml.aabbCenter[0], ml.aabbCenter[1], ml.aabbCenter[2]};
		const float w3[3] = {we.x, we.y, we.z};
		for (int r = 0; r < 3; ++r)
			posErr[r] = kNoise * (((std::fabs(T[r] * c[0]) + std::fabs(T[4 + r] * c[1])) + std::fabs(T[8 + r] * c[2])) + std::fabs(T[12 + r]) + w3[r]);
	}

	// :52 -> culling.h.glsl:8-19
	for (int i = 0; i < 6; ++i) {
		const float* p = cam.frustum[i];
		float radius = dot3(we, V3{std::fabs(p[0]), std::fabs(p[1]), std::fabs(p[2])});
		float distance = dot3(V3{p[0], p[1], p[2]}, wc) - p[3];
		if (diag) {
			const float terms = ((std::fabs(p[0] * wc.x) + std::fabs(p[1] * wc.y)) + std::fabs(p[2] * wc.z)) + std::fabs(p[3]) + std::fabs(radius);
			const float moved = (std::fabs(p[0]) * posErr[0] + std::fabs(p[1]) * posErr[1]) + std::fabs(p[2]) * posErr[2];
			if (std::fabs(-radius - distance) <= kNoise * terms + moved) flags |= ORC_AMBIG_FRUSTUM;
		}
		if (-radius > distance) return ORC_FRUSTUM_CULLED | flags;
	}

	if (cone) { // dot(apex - eye, axis) >= (cutoff + margin) * |apex - eye|   (meshoptimizer.h:531 without the normalisation)
		const vkv_MeshletCone& cn = ((const vkv_MeshletCone*)cone->table[d.primitiveIndex])[d.meshletIndex];
		const V4 eye = cone->eye[d.transformIndex];
		const float dx = cn.apex[0] - eye.x, dy = cn.apex[1] - eye.y, dz = cn.apex[2] - eye.z;
		const float len2 = (dx * dx + dy * dy) + dz * dz;
		const float dp = (dx * cn.axis[0] + dy * cn.axis[1]) + dz * cn.axis[2];
		if (dp >= (cn.cutoff + kConeMargin) * std::sqrt(len2)) return ORC_FRUSTUM_CULLED | ORC_CONE_CULLED | flags;
	}

	// :56 -> culling.h.glsl:44-56
	V3 ssMin{1.f, 1.f, 1.f}, ssMax{-1.f, -1.f, -1.f};
	float uvErr = 0.f, zErr = 0.f;   // noise of any corner's uv / depth (first order)
	for (int i = 0; i < 8; ++i) {
		V4 pos{kAabbPositions[i][0] * we.x + wc.x, kAabbPositions[i][1] * we.y + wc.y, kAabbPositions[i][2] * we.z + wc.z, 1.0f};
		V4 clip = mul44(occVP, pos);
		if (!(clip.w > 0.0f)) flags |= ORC_CROSSES_CAMERA;
		float ndcx = gclamp(clip.x / clip.w, -1.f, 1.f);
		float ndcy = gclamp(clip.y / clip.w, -1.f, 1.f);
		float uvx = ndcx * 0.5f + 0.5f;
		float uvy = ndcy * 0.5f + 0.5f;
		float z = clip.z / clip.w;
		if (diag) {
			float cerr[4];
			for (int r = 0; r < 4; ++r) {
				const float terms = ((std::fabs(occVP[r] * pos.x) + std::fabs(occVP[4 + r] * pos.y)) + std::fabs(occVP[8 + r] * pos.z)) + std::fabs(occVP[12 + r]);
				const float moved = (std::fabs(occVP[r]) * posErr[0] + std::fabs(occVP[4 + r]) * posErr[1]) + std::fabs(occVP[8 + r]) * posErr[2];
				cerr[r] = kNoise * terms + moved;
			}
			const float iw = 1.0f / std::fabs(clip.w);
			const float ex = (cerr[0] + std::fabs(clip.x * iw) * cerr[3]) * iw, ey = (cerr[1] + std::fabs(clip.y * iw) * cerr[3]) * iw;
			const float ez = (cerr[2] + std::fabs(z) * cerr[3]) * iw + kNoise * std::fabs(z);
			uvErr = std::fmax(uvErr, 0.5f * std::fmax(ex, ey) + kNoise);
			zErr = std::fmax(zErr, ez);
		}
		ssMin.x = gmin(ssMin.x, uvx); ssMin.y = gmin(ssMin.y, uvy); ssMin.z = gmin(ssMin.z, z);
		ssMax.x = gmax(ssMax.x, uvx); ssMax.y = gmax(ssMax.y, uvy); ssMax.z = gmax(ssMax.z, z);
	}
	// :57-59 ; pyramidSize = textureSize(pyramid, 0) (:33)
	float width = (ssMax.x - ssMin.x) * (float)(int)pyr.w[0];
	float height = (ssMax.y - ssMin.y) * (float)(int)pyr.h[0];
	float m = gmax(width, height);
	// floor(log2(m)) as the exact binary exponent; the sampler clamps lod to [minLod,maxLod]=[0,16] and to
	// the existing mips (application.cpp:451-452).  NaN / <=0 -> level 0 ; +inf -> last mip.
	auto level_of = [&](float v) {
		int l;
		if (!(v > 0.0f)) l = 0;
		else if (std::isinf(v)) l = 16;
		else { l = std::ilogb(v); if (l < 0) l = 0; if (l > 16) l = 16; }
		if (l > (int)pyr.levels - 1) l = (int)pyr.levels - 1;
		return l;
	};
	const int level = level_of(m);
	if (diag) {
		// SURVEY Q6 (an implementation's log2 may round up just below a power of two) and the noise of the extent itself
		const float mErr = 2.0f * uvErr * (float)(int)std::max(pyr.w[0], pyr.h[0]) + 2.0f * 1.1920929e-7f * std::fabs(m);
		if (level_of(m + mErr) != level || level_of(m - mErr) != level) flags |= ORC_AMBIG_LEVEL;
	}
	// :61-62
	float cx = (ssMin.x + ssMax.x) * 0.5f;
	float cy = (ssMin.y + ssMax.y) * 0.5f;
	int amb = 0;
	float depth = sample_min(pyramid + pyr.off[level], pyr.w[level], pyr.h[level], cx, cy, &amb);
	if (diag) {
		// the footprint {floor(u), floor(u)+1} changes when u = c*size - 0.5 moves across an integer
		auto crosses = [&](float c, uint32_t size) {
			const float u = c * (float)size - 0.5f, du = uvErr * (float)size + 4.0f * 1.1920929e-7f * std::fabs(u) + 1e-6f;
			return std::floor(u - du) != std::floor(u + du) || u - std::floor(u) == 0.0f;
		};
		if (amb || crosses(cx, pyr.w[level]) || crosses(cy, pyr.h[level])) flags |= ORC_AMBIG_FOOTPRINT;
		if (std::fabs(depth - ssMax.z) <= zErr + 4.0f * 1.1920929e-7f * std::fmax(std::fabs(depth), std::fabs(ssMax.z))) flags |= ORC_AMBIG_HIZ;
	}
	// :64
	bool visible = depth < ssMax.z;
	return (visible ? ORC_VISIBLE : ORC_OCCLUDED) | flags;
}

// ---------------------------------------------------------------------------------------------
// Rasteriser.  Fixed-function rules restated:
//   viewport (0,0,W,H), depth [0,1]; pixel centres at +0.5; 8 sub-pixel bits; top-left fill rule;
//   cullMode NONE; depthClamp off => clip to 0<=z<=w; guard band |x|,|y| <= 8w (implementation choice);
//   depth = screen-space linear interpolation of z/w; test GREATER_OR_EQUAL, clear 0.0.
// ---------------------------------------------------------------------------------------------
constexpr int kSubBits = 8;
constexpr int kSub = 1 << kSubBits;
constexpr float kGuard = 8.0f;

struct SetupTri {
	int32_t ax, ay, bx, by, cx, cy;
	int32_t xmin, xmax, ymin, ymax;
	int64_t area2;
	float za, dzb, dzc, invA;
	uint32_t id;
};

struct RasterStats {
	uint64_t tris_in = 0, facing = 0, rejected = 0, clipped = 0, degenerate = 0, rasterised = 0;
};

inline bool has_nan(const V4& v) { return !(v.x == v.x && v.y == v.y && v.z == v.z && v.w == v.w); }

// One already-inside (post-clip) triangle -> SetupTri
inline void setup_projected(const V4& A, const V4& B, const V4& C, uint32_t id, uint32_t W, uint32_t H,
                            std::vector<SetupTri>& out, RasterStats& st) {
	if (!(A.w > 0.f) || !(B.w > 0.f) || !(C.w > 0.f)) { st.degenerate++; return; }
	const float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
	auto proj = [&](const V4& v, int32_t& fx, int32_t& fy, float& z) {
		float nx = v.x / v.w, ny = v.y / v.w;
		z = v.z / v.w;
		float sx = nx * hw + hw;
		float sy = ny * hh + hh;
		fx = (int32_t)std::lrintf(sx * (float)kSub);
		fy = (int32_t)std::lrintf(sy * (float)kSub);
	};
	SetupTri t;
	float za, zb, zc;
	proj(A, t.ax, t.ay, za);
	proj(B, t.bx, t.by, zb);
	proj(C, t.cx, t.cy, zc);
	int64_t area2 = (int64_t)(t.bx - t.ax) * (t.cy - t.ay) - (int64_t)(t.by - t.ay) * (t.cx - t.ax);
	if (area2 == 0) { st.degenerate++; return; }
	if (area2 < 0) { // cullMode NONE: both windings are drawn; normalise to positive area
		std::swap(t.bx, t.cx); std::swap(t.by, t.cy); std::swap(zb, zc);
		area2 = -area2;
	}
	t.area2 = area2;
	t.za = za; t.dzb = zb - za; t.dzc = zc - za;
	t.invA = 1.0f / (float)area2;
	int32_t minx = std::min(t.ax, std::min(t.bx, t.cx)), maxx = std::max(t.ax, std::max(t.bx, t.cx));
	int32_t miny = std::min(t.ay, std::min(t.by, t.cy)), maxy = std::max(t.ay, std::max(t.by, t.cy));
	t.xmin = std::max<int32_t>(0, (minx + (kSub / 2 - 1)) >> kSubBits);
	t.xmax = std::min<int32_t>((int32_t)W - 1, (maxx - kSub / 2) >> kSubBits);
	t.ymin = std::max<int32_t>(0, (miny + (kSub / 2 - 1)) >> kSubBits);
	t.ymax = std::min<int32_t>((int32_t)H - 1, (maxy - kSub / 2) >> kSubBits);
	if (t.xmin > t.xmax || t.ymin > t.ymax) { st.rejected++; return; }
	t.id = id;
	out.push_back(t);
	st.rasterised++;
}

// Sutherland–Hodgman against one plane; dist(v) >= 0 is inside.  The intersection is always evaluated
// from the inside vertex towards the outside vertex so a shared edge yields the same point in both triangles.
template <class D>
inline int clip_plane(const V4* in, int n, V4* out, D dist) {
	int m = 0;
	for (int i = 0; i < n; ++i) {
		const V4& cur = in[i];
		const V4& nxt = in[(i + 1) % n];
		float dc = dist(cur), dn = dist(nxt);
		bool ic = dc >= 0.f, inx = dn >= 0.f;
		if (ic) out[m++] = cur;
		if (ic != inx) {
			const V4& a = ic ? cur : nxt; // inside
			const V4& b = ic ? nxt : cur; // outside
			float da = ic ? dc : dn, db = ic ? dn : dc;
			float t = da / (da - db);
			V4 r;
			r.x = a.x + t * (b.x - a.x);
			r.y = a.y + t * (b.y - a.y);
			r.z = a.z + t * (b.z - a.z);
			r.w = a.w + t * (b.w - a.w);
			out[m++] = r;
		}
	}
	return m;
}

inline void raster_setup(const V4& A, const V4& B, const V4& C, uint32_t id, uint32_t W, uint32_t H,
                         std::vector<SetupTri>& out, RasterStats& st) {
	if (has_nan(A) || has_nan(B) || has_nan(C)) { st.rejected++; return; }
	// trivial reject: all three vertices outside one view-volume plane
	auto code = [](const V4& v) {
		int c = 0;
		if (v.x < -v.w) c |= 1;
		if (v.x > v.w) c |= 2;
		if (v.y < -v.w) c |= 4;
		if (v.y > v.w) c |= 8;
		if (v.z < 0.f) c |= 16;
		if (v.z > v.w) c |= 32;
		return c;
	};
	int ca = code(A), cb = code(B), cc = code(C);
	if (ca & cb & cc) { st.rejected++; return; }
	auto needs = [](const V4& v) {
		float g = kGuard * v.w;
		return v.z < 0.f || v.z > v.w || v.x > g || v.x < -g || v.y > g || v.y < -g || !(v.w > 0.f);
	};
	if (!(needs(A) || needs(B) || needs(C))) { setup_projected(A, B, C, id, W, H, out, st); return; }
	st.clipped++;
	V4 p0[12], p1[12];
	p0[0] = A; p0[1] = B; p0[2] = C;
	int n = 3;
	n = clip_plane(p0, n, p1, [](const V4& v) { return v.w - v.z; }); if (n < 3) return;      // near (reverse-Z: z<=w)
	n = clip_plane(p1, n, p0, [](const V4& v) { return v.z; }); if (n < 3) return;            // far  (z>=0)
	n = clip_plane(p0, n, p1, [](const V4& v) { return kGuard * v.w - v.x; }); if (n < 3) return;
	n = clip_plane(p1, n, p0, [](const V4& v) { return kGuard * v.w + v.x; }); if (n < 3) return;
	n = clip_plane(p0, n, p1, [](const V4& v) { return kGuard * v.w - v.y; }); if (n < 3) return;
	n = clip_plane(p1, n, p0, [](const V4& v) { return kGuard * v.w + v.y; }); if (n < 3) return;
	for (int i = 1; i + 1 < n; ++i) setup_projected(p0[0], p0[i], p0[i + 1], id, W, H, out, st);
}

// visbuffer.mesh.glsl:30-104 for one surviving MeshletDraw -> setup triangles
void meshlet_setup(const vkv_VisbufferPushConstants* pc, uint32_t drawId, const vkv_Camera& cam, uint32_t W, uint32_t H,
                   std::vector<SetupTri>& out, RasterStats& st) {
	const vkv_MeshletDraw d = ((const vkv_MeshletDraw*)pc->drawBuffer)[drawId];           // mesh.glsl:32
	const vkv_Primitive& prim = ((const vkv_Primitive*)pc->primitiveBuffer)[d.primitiveIndex];
	const vkv_Meshlet& ml = ((const vkv_Meshlet*)prim.meshletBuffer)[d.meshletIndex];
	const vkv_Material& mat = ((const vkv_Material*)pc->materialBuffer)[prim.materialIndex];
	const float* T = (const float*)pc->transformBuffer + (size_t)d.transformIndex * 16;  // :43
	float mvp[16];
	mul44m(cam.viewProjection, T, mvp);                                                  // :44
	const uint32_t* vidx = (const uint32_t*)prim.vertexIndexBuffer;
	const vkv_Vertex* verts = (const vkv_Vertex*)prim.vertexBuffer;
	const uint8_t* tris = (const uint8_t*)prim.primitiveIndexBuffer;
	V4 clip[VKV_MAX_VERTICES];
	uint32_t vc = ml.vertexCount, tc = ml.triangleCount;
	for (uint32_t v = 0; v < vc && v < VKV_MAX_VERTICES; ++v) {                          // :50-69
		const float* p = verts[vidx[ml.vertexOffset + v]].position;
		clip[v] = mul44(mvp, V4{p[0], p[1], p[2], 1.0f});                                // :61
	}
	float transformDet = det4(T);                                                        // :71
	bool doubleSided = mat.doubleSided != 0;
	for (uint32_t t = 0; t < tc; ++t) {                                                  // :73-103
		uint32_t a = tris[ml.triangleOffset + t * 3 + 0], b = tris[ml.triangleOffset + t * 3 + 1],
		         c = tris[ml.triangleOffset + t * 3 + 2];
		st.tris_in++;
		if (!doubleSided) {                                                              // :86-98
			float det = det3(V3{clip[a].x, clip[a].y, clip[a].w}, V3{clip[b].x, clip[b].y, clip[b].w},
			                 V3{clip[c].x, clip[c].y, clip[c].w});
			bool cull = (transformDet < 0.0f) ? (det < 0.0f) : (det > 0.0f);
			if (cull) { st.facing++; continue; }
		}
		raster_setup(clip[a], clip[b], clip[c], vkv_pack_visbuffer(drawId, t), W, H, out, st);   // frag.glsl:36
	}
}

inline bool top_left(int32_t dx, int32_t dy) { return dy < 0 || (dy == 0 && dx > 0); }

// rasterise rows [y0,y1) of one set-up triangle
inline void raster_rows(const SetupTri& t, int y0, int y1, uint32_t W, float* depth, uint32_t* ids_ref, uint32_t* ids_min,
                        uint8_t* tie, uint64_t& frags, uint64_t& passed) {
	int ys = std::max(y0, t.ymin), ye = std::min(y1 - 1, t.ymax);
	if (ys > ye) return;
	// edge k opposite vertex k: w0 = E(b,c,p), w1 = E(c,a,p), w2 = E(a,b,p); E(u,v,p) = (vx-ux)(py-uy) - (vy-uy)(px-ux)
	const int64_t e0dx = t.cx - t.bx, e0dy = t.cy - t.by;
	const int64_t e1dx = t.ax - t.cx, e1dy = t.ay - t.cy;
	const int64_t e2dx = t.bx - t.ax, e2dy = t.by - t.ay;
	const int64_t b0 = top_left((int32_t)e0dx, (int32_t)e0dy) ? 0 : 1;
	const int64_t b1 = top_left((int32_t)e1dx, (int32_t)e1dy) ? 0 : 1;
	const int64_t b2 = top_left((int32_t)e2dx, (int32_t)e2dy) ? 0 : 1;
	for (int y = ys; y <= ye; ++y) {
		const int64_t py = (int64_t)y * kSub + kSub / 2;
		const int64_t px0 = (int64_t)t.xmin * kSub + kSub / 2;
		int64_t w0 = e0dx * (py - t.by) - e0dy * (px0 - t.bx);
		int64_t w1 = e1dx * (py - t.cy) - e1dy * (px0 - t.cx);
		int64_t w2 = e2dx * (py - t.ay) - e2dy * (px0 - t.ax);
		size_t row = (size_t)y * W;
		for (int x = t.xmin; x <= t.xmax; ++x) {
			if (w0 >= b0 && w1 >= b1 && w2 >= b2) {
				float l1 = (float)w1 * t.invA;
				float l2 = (float)w2 * t.invA;
				float z = (t.za + l1 * t.dzb) + l2 * t.dzc;
				z = (z > 0.0f) ? z : 0.0f;
				z = (z < 1.0f) ? z : 1.0f;
				frags++;
				size_t p = row + x;
				float d = depth[p];
				if (z > d) {                       // GREATER_OR_EQUAL, strict part
					depth[p] = z; ids_ref[p] = t.id; ids_min[p] = t.id; tie[p] = 0; passed++;
				} else if (z == d) {               // GREATER_OR_EQUAL, equal part: the later primitive wins
					if (ids_ref[p] == VKV_VISBUFFER_CLEAR) { ids_ref[p] = t.id; ids_min[p] = t.id; tie[p] = 0; }
					else {
						if (t.id != ids_min[p]) { tie[p] = 1; if (t.id < ids_min[p]) ids_min[p] = t.id; }
						ids_ref[p] = t.id;
					}
					passed++;
				}
			}
			w0 -= e0dy * kSub; w1 -= e1dy * kSub; w2 -= e2dy * kSub;
		}
	}
}

} // namespace

extern "C" {

uint32_t orc_pyramid_layout(uint32_t W, uint32_t H, uint32_t offsets[17], uint32_t w[16], uint32_t h[16], uint32_t* total) {
	uint32_t levels = vkv_mip_levels(W, H);
	if (levels > 16) levels = 16;
	uint32_t off = 0;
	for (uint32_t k = 0; k < levels; ++k) {
		w[k] = vkv_mip_extent(W, k);
		h[k] = vkv_mip_extent(H, k);
		offsets[k] = off;
		off += w[k] * h[k];
	}
	offsets[levels] = off;
	if (total) *total = off;
	return levels;
}

void orc_set_diagnostics(int on) { g_diagnostics = on != 0; }

float orc_sample_min(const float* img, uint32_t w, uint32_t h, float u, float v, int* ambig) {
	return sample_min(img, w, h, u, v, ambig);
}

uint64_t orc_vis64_key(float depth, uint32_t id) { return ((uint64_t)(~fbits(depth)) << 32) | id; }

int orc_cull(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const float* pyramid, int vp_select,
             const uint8_t* only_status, uint8_t* status, orc_counters* ctr, int threads) {
	return orc_cull_cone(pc, W, H, pyramid, vp_select, only_status, status, ctr, threads, nullptr);
}

int orc_cull_cone(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const float* pyramid, int vp_select,
                  const uint8_t* only_status, uint8_t* status, orc_counters* ctr, int threads, const uint64_t* cone_table) {
	Pyr pyr;
	pyr.levels = orc_pyramid_layout(W, H, pyr.off, pyr.w, pyr.h, &pyr.total);
	if (pyr.levels == 0) return -1;
	const vkv_Camera cam = *(const vkv_Camera*)pc->cameraBuffer; // task.glsl:31
	const float* occVP = vp_select == 0 ? cam.prevOcclusionViewProjection : cam.viewProjection;
	const uint32_t N = pc->meshletDrawCount;
	std::vector<V4> eyes;
	ConeCtx coneCtx{cone_table, nullptr};
	if (cone_table) { // the current camera's position per mesh node (the raster that follows uses viewProjection, whichever VP the HiZ test uses)
		const vkv_MeshletDraw* draws = (const vkv_MeshletDraw*)pc->drawBuffer;
		uint32_t nT = 0;
		for (uint32_t i = 0; i < N; ++i) nT = std::max(nT, draws[i].transformIndex + 1);
		eyes.resize(nT);
		for (uint32_t t = 0; t < nT; ++t) eyes[t] = mesh_space_eye(cam.viewProjection, (const float*)pc->transformBuffer + (size_t)t * 16);
		coneCtx.eye = eyes.data();
	}
	const ConeCtx* cone = cone_table ? &coneCtx : nullptr;
	// chunks of maxMeshlets draws = one reference task workgroup (task.glsl:28-29)
	const size_t chunks = (N + VKV_MAX_MESHLETS_PER_TASK - 1) / VKV_MAX_MESHLETS_PER_TASK;
	parallel_for(chunks, threads, [&](size_t b, size_t e, int) {
		for (size_t c = b; c < e; ++c) {
			uint32_t lo = (uint32_t)c * VKV_MAX_MESHLETS_PER_TASK, hi = std::min<uint32_t>(N, lo + VKV_MAX_MESHLETS_PER_TASK);
			for (uint32_t i = lo; i < hi; ++i) {
				if (only_status && (only_status[i] & ORC_STATUS_MASK) != ORC_OCCLUDED) { status[i] = ORC_NOT_TESTED; continue; }
				status[i] = cull_one(pc, i, cam, occVP, pyr, pyramid, cone);
			}
		}
	});
	if (ctr) {
		for (uint32_t i = 0; i < N; ++i) {
			uint8_t s = status[i];
			if ((s & ORC_STATUS_MASK) == ORC_NOT_TESTED) continue;
			ctr->tested++;
			switch (s & ORC_STATUS_MASK) {
				case ORC_FRUSTUM_CULLED: ctr->frustum_culled++; break;
				case ORC_OCCLUDED: ctr->occluded++; break;
				case ORC_VISIBLE: ctr->visible++; break;
			}
			if (s & ORC_AMBIG_FRUSTUM) ctr->ambig_frustum++;
			if (s & ORC_AMBIG_HIZ) ctr->ambig_hiz++;
			if (s & ORC_AMBIG_LEVEL) ctr->ambig_level++;
			if (s & ORC_AMBIG_FOOTPRINT) ctr->ambig_footprint++;
			if (s & ORC_CROSSES_CAMERA) ctr->crosses_camera++;
		}
	}
	return 0;
}

void orc_clear(uint32_t W, uint32_t H, float* depth, uint32_t* ids_ref, uint32_t* ids_min, uint8_t* tie) {
	size_t n = (size_t)W * H;
	for (size_t i = 0; i < n; ++i) {
		depth[i] = 0.0f;                                   // application.cpp:807
		if (ids_ref) ids_ref[i] = VKV_VISBUFFER_CLEAR;     // application.cpp:782
		if (ids_min) ids_min[i] = VKV_VISBUFFER_CLEAR;
		if (tie) tie[i] = 0;
	}
}

int orc_raster(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const uint32_t* draw_ids, uint32_t n_draws,
               float* depth, uint32_t* ids_ref, uint32_t* ids_min, uint8_t* tie, orc_counters* ctr, int threads) {
	if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
	if (threads < 1) threads = 1;
	const vkv_Camera cam = *(const vkv_Camera*)pc->cameraBuffer;
	// Phase 1: geometry + triangle setup, parallel over contiguous chunks of the draw list (order preserved).  Every chunk also bins
	// its triangles by the horizontal bands of phase 2 (indices in submission order), so a band walks only what can touch it.
	const int nchunks = std::max(1, std::min<int>(threads * 4, (int)n_draws));
	const int bands = std::max(1, std::min<int>(threads * 8, (int)H));
	std::vector<uint16_t> band_of(H);
	for (int b = 0; b < bands; ++b)
		for (size_t y = (size_t)H * b / bands, e = (size_t)H * (b + 1) / bands; y < e; ++y) band_of[y] = (uint16_t)b;
	std::vector<std::vector<SetupTri>> lists(nchunks);
	std::vector<std::vector<std::vector<uint32_t>>> bins(nchunks, std::vector<std::vector<uint32_t>>(bands));
	std::vector<RasterStats> stats(nchunks);
	std::atomic<int> next{0};
	auto work1 = [&] {
		for (;;) {
			int c = next.fetch_add(1);
			if (c >= nchunks) break;
			size_t b = (size_t)n_draws * c / nchunks, e = (size_t)n_draws * (c + 1) / nchunks;
			std::vector<SetupTri>& L = lists[c];
			for (size_t i = b; i < e; ++i) {
				const size_t first = L.size();
				meshlet_setup(pc, draw_ids[i], cam, W, H, L, stats[c]);
				for (size_t k = first; k < L.size(); ++k) {
					const int ya = std::max(L[k].ymin, 0), yb = std::min(L[k].ymax, (int)H - 1);
					if (ya > yb) continue;
					for (int bb = band_of[ya]; bb <= band_of[yb]; ++bb) bins[c][bb].push_back((uint32_t)k);
				}
			}
		}
	};
	{
		std::vector<std::thread> pool;
		for (int t = 1; t < threads; ++t) pool.emplace_back(work1);
		work1();
		for (auto& th : pool) th.join();
	}
	// Phase 2: pixels, parallel over horizontal bands; each band walks its triangles in submission order (chunk by chunk, index by
	// index), so the result equals a sequential rasteriser.
	std::vector<uint64_t> frags(bands, 0), passed(bands, 0);
	std::atomic<int> nextb{0};
	auto work2 = [&] {
		for (;;) {
			int b = nextb.fetch_add(1);
			if (b >= bands) break;
			int y0 = (int)((size_t)H * b / bands), y1 = (int)((size_t)H * (b + 1) / bands);
			for (int c = 0; c < nchunks; ++c)
				for (uint32_t k : bins[c][b]) raster_rows(lists[c][k], y0, y1, W, depth, ids_ref, ids_min, tie, frags[b], passed[b]);
		}
	};
	{
		std::vector<std::thread> pool;
		for (int t = 1; t < threads; ++t) pool.emplace_back(work2);
		work2();
		for (auto& th : pool) th.join();
	}
	if (ctr) {
		ctr->meshlets += n_draws;
		for (auto& s : stats) {
			ctr->triangles_in += s.tris_in; ctr->triangles_culled_facing += s.facing; ctr->triangles_rejected += s.rejected;
			ctr->triangles_clipped += s.clipped; ctr->triangles_degenerate += s.degenerate; ctr->triangles_rasterised += s.rasterised;
		}
		for (int b = 0; b < bands; ++b) { ctr->fragments += frags[b]; ctr->fragments_passed += passed[b]; }
		uint64_t ties = 0;
		for (size_t i = 0, n = (size_t)W * H; i < n; ++i) ties += tie[i];
		ctr->tie_pixels = ties;
	}
	return 0;
}

// The mesh shader's per-vertex / per-triangle results for the given MeshletDraws, in the layout of ref_shim.cpp's ref_mesh_shader
// (the same lines of visbuffer.mesh.glsl evaluated by the reference's text against glm), for tests/test_oracle.py:
//   clip[d][v]   gl_Position of meshlet vertex v (:61)               64 slots x 4 floats
//   cull[d][t]   gl_CullPrimitiveEXT of triangle t (:86-102)         126 slots; 0 / 1, 0xff = beyond triangleCount
//   det[d][t]    determinant(mat3(v0, v1, v2)) (:93)                 the oracle's association (det3 above)
//   tdet[d]      determinant(transformMatrix) (:71)
//   noise[d][t]  first-order bound on how far an equally valid evaluation of the same GLSL can move det: kNoise roundings of the largest
//                term of every sum (mat4*mat4, mat4*vec4, the six triple products), propagated through the determinant's cofactors
//   ambig[d][t]  1 = |det| <= noise (or |transformDet| within its own noise): such an evaluation could decide the triangle differently
void orc_mesh_shader(const vkv_VisbufferPushConstants* pc, const uint32_t* draw_ids, uint32_t n, float* clip, uint8_t* cull, float* det,
                     float* tdet, uint8_t* ambig, float* noise) {
	const vkv_Camera cam = *(const vkv_Camera*)pc->cameraBuffer;
	constexpr uint32_t kSlots = 126; // mesh_common.h.glsl maxPrimitives
	for (uint32_t d = 0; d < n; ++d) {
		const vkv_MeshletDraw dr = ((const vkv_MeshletDraw*)pc->drawBuffer)[draw_ids[d]];
		const vkv_Primitive& prim = ((const vkv_Primitive*)pc->primitiveBuffer)[dr.primitiveIndex];
		const vkv_Meshlet& ml = ((const vkv_Meshlet*)prim.meshletBuffer)[dr.meshletIndex];
		const vkv_Material& mat = ((const vkv_Material*)pc->materialBuffer)[prim.materialIndex];
		const float* T = (const float*)pc->transformBuffer + (size_t)dr.transformIndex * 16;
		float mvp[16], mvpErr[16];
		mul44m(cam.viewProjection, T, mvp);
		for (int c = 0; c < 4; ++c)
			for (int r = 0; r < 4; ++r) {
				float terms = 0.f;
				for (int k = 0; k < 4; ++k) terms += std::fabs(cam.viewProjection[k * 4 + r] * T[c * 4 + k]);
				mvpErr[c * 4 + r] = kNoise * terms;
			}
		const uint32_t* vidx = (const uint32_t*)prim.vertexIndexBuffer;
		const vkv_Vertex* verts = (const vkv_Vertex*)prim.vertexBuffer;
		const uint8_t* tris = (const uint8_t*)prim.primitiveIndexBuffer;
		V4 cv[VKV_MAX_VERTICES];
		float cErr[VKV_MAX_VERTICES][4];
		const uint32_t vc = ml.vertexCount < VKV_MAX_VERTICES ? ml.vertexCount : VKV_MAX_VERTICES;
		for (uint32_t v = 0; v < vc; ++v) {
			const float* p = verts[vidx[ml.vertexOffset + v]].position;
			cv[v] = mul44(mvp, V4{p[0], p[1], p[2], 1.0f});
			std::memcpy(clip + ((size_t)d * VKV_MAX_VERTICES + v) * 4, &cv[v], 16);
			for (int r = 0; r < 4; ++r) {
				const float terms = ((std::fabs(mvp[r] * p[0]) + std::fabs(mvp[4 + r] * p[1])) + std::fabs(mvp[8 + r] * p[2])) + std::fabs(mvp[12 + r]);
				const float moved = ((mvpErr[r] * std::fabs(p[0]) + mvpErr[4 + r] * std::fabs(p[1])) + mvpErr[8 + r] * std::fabs(p[2])) + mvpErr[12 + r];
				cErr[v][r] = kNoise * terms + moved;
			}
		}
		const float transformDet = det4(T);
		tdet[d] = transformDet;
		// noise of det4: the cofactor sum's terms are at most |m0 d0| + ... ; each d_k is itself a sum of six triple products
		float tdetTerms = 0.f;
		{
			float a[16];
			for (int i = 0; i < 16; ++i) a[i] = std::fabs(T[i]);
			// permanent of |T| bounds the sum of the absolute values of all 24 products
			auto perm3 = [&](int c0, int c1, int c2, int skipRow) {
				int rows[3], k = 0;
				for (int r = 0; r < 4; ++r) if (r != skipRow) rows[k++] = r;
				auto A = [&](int c, int r) { return a[c * 4 + rows[r]]; };
				return A(c0, 0) * (A(c1, 1) * A(c2, 2) + A(c1, 2) * A(c2, 1)) + A(c0, 1) * (A(c1, 0) * A(c2, 2) + A(c1, 2) * A(c2, 0)) +
				       A(c0, 2) * (A(c1, 0) * A(c2, 1) + A(c1, 1) * A(c2, 0));
			};
			for (int r = 0; r < 4; ++r) tdetTerms += a[r] * perm3(1, 2, 3, r);
		}
		const bool tdetAmbig = std::fabs(transformDet) <= 2.0f * kNoise * tdetTerms;
		for (uint32_t t = 0; t < kSlots; ++t) {
			uint8_t& c = cull[(size_t)d * kSlots + t];
			c = 0xff; det[(size_t)d * kSlots + t] = 0.f; ambig[(size_t)d * kSlots + t] = 0; noise[(size_t)d * kSlots + t] = 0.f;
			if (t >= ml.triangleCount) continue;
			const uint32_t ia = tris[ml.triangleOffset + t * 3 + 0], ib = tris[ml.triangleOffset + t * 3 + 1], ic = tris[ml.triangleOffset + t * 3 + 2];
			if (mat.doubleSided != 0 || ia >= vc || ib >= vc || ic >= vc) { c = 0; continue; }
			const V3 a{cv[ia].x, cv[ia].y, cv[ia].w}, b{cv[ib].x, cv[ib].y, cv[ib].w}, cc{cv[ic].x, cv[ic].y, cv[ic].w};
			const float dt = det3(a, b, cc);
			det[(size_t)d * kSlots + t] = dt;
			c = ((transformDet < 0.0f) ? (dt < 0.0f) : (dt > 0.0f)) ? 1 : 0;
			// first-order bound: |d det| <= sum_i |cofactor_i| * err_i  +  roundings of the six triple products
			auto absv = [](V3 v) { return V3{std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)}; };
			auto permCross = [](V3 u, V3 v) { return V3{u.y * v.z + u.z * v.y, u.z * v.x + u.x * v.z, u.x * v.y + u.y * v.x}; }; // |cross| bound
			const V3 A = absv(a), B = absv(b), C = absv(cc);
			const V3 eA{cErr[ia][0], cErr[ia][1], cErr[ia][3]}, eB{cErr[ib][0], cErr[ib][1], cErr[ib][3]}, eC{cErr[ic][0], cErr[ic][1], cErr[ic][3]};
			const float moved = dot3(eA, permCross(B, C)) + dot3(eB, permCross(A, C)) + dot3(eC, permCross(A, B));
			const float terms = dot3(A, permCross(B, C)); // sum of the absolute values of the six triple products
			const float bound = 2.0f * kNoise * terms + 2.0f * moved;
			noise[(size_t)d * kSlots + t] = bound;
			if (tdetAmbig || std::fabs(dt) <= bound) ambig[(size_t)d * kSlots + t] = 1;
		}
	}
}

// srgb.h.glsl:26-32 (per channel) and the RGBA8_UNORM image store conversion
static float from_linear(float c) {
	bool cutoff = c < 0.0031308f;
	float higher = 1.055f * std::pow(c, 1.f / 2.4f) - 0.055f;
	float lower = c * 12.92f;
	return cutoff ? lower : higher;
}
static uint32_t unorm8(float v) {
	v = (v > 0.0f) ? v : 0.0f;
	v = (v < 1.0f) ? v : 1.0f;
	return (uint32_t)std::lrintf(v * 255.0f);
}

float orc_from_linear(float c) { return from_linear(c); } // for the pin against the reference's srgb.h.glsl (tests/test_oracle.py)

int orc_resolve(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const uint32_t* ids, uint32_t* out) {
	if (pc->meshletDrawCount == 0) return 0;                       // application.cpp:930
	const vkv_MeshletDraw* draws = (const vkv_MeshletDraw*)pc->drawBuffer;
	const vkv_Primitive* prims = (const vkv_Primitive*)pc->primitiveBuffer;
	const vkv_Material* mats = (const vkv_Material*)pc->materialBuffer;
	const uint32_t covered = (W / 32u) * 32u;                      // application.cpp:943
	for (uint32_t y = 0; y < H; ++y)
		for (uint32_t x = 0; x < covered; ++x) {
			size_t p = (size_t)y * W + x;
			out[p] = 0;                                            // comp.glsl:25
			uint32_t v = ids[p];
			if (v == VKV_VISBUFFER_CLEAR) continue;                // :28
			uint32_t drawIndex = v >> VKV_TRIANGLE_BITS;           // :33 unpackVisBuffer
			const vkv_Material& m = mats[prims[draws[drawIndex].primitiveIndex].materialIndex];   // :35-37
			out[p] = unorm8(from_linear(m.albedoFactor[0])) | (unorm8(from_linear(m.albedoFactor[1])) << 8) |
			         (unorm8(from_linear(m.albedoFactor[2])) << 16) | (unorm8(m.albedoFactor[3]) << 24);   // :39-41
		}
	return 0;
}

// fp32 -> fp16 with round-to-nearest-even (the R16G16_SFLOAT store of the motion-vector attachment), through exact double scaling and
// nearbyint (the default rounding mode is to-nearest-even); NaN -> 0x7FFF
static uint16_t to_half(float f) {
	if (f != f) return 0x7fffu;
	const uint16_t sign = std::signbit(f) ? 0x8000u : 0u;
	const double a = std::fabs((double)f);
	if (a >= 65520.0) return (uint16_t)(sign | 0x7c00u);                       // the tie between 65504 and 2^16 goes to even = infinity
	if (a < 6.103515625e-05) return (uint16_t)(sign | (uint16_t)std::nearbyint(a * 16777216.0)); // below 2^-14: units of 2^-24 (1024 = 2^-14)
	int e;
	const double m = std::frexp(a, &e);                                        // a = m * 2^e, m in [0.5, 1)
	long q = std::lrint(m * 2048.0);                                           // 11 significant bits, 1024 .. 2048
	if (q == 2048) { q = 1024; ++e; }
	return (uint16_t)(sign | (uint16_t)(((e + 14) << 10) + (q - 1024)));
}
uint16_t orc_to_half(float f) { return to_half(f); }

// The visbuffer pass's second colour attachment (application.cpp:250-267 R16G16_SFLOAT, cleared to 0 at :786-799): visbuffer.frag.glsl:38
// with the varyings of visbuffer.mesh.glsl:44-45,61-63, evaluated once per pixel for the triangle the id image names (what the depth test
// left visible).  Perspective-correct interpolation at the pixel centre through homogeneous edge functions in (x, y, w); the common divisor of
// the interpolated varyings cancels in the shader's ratios.  out_f: 2 floats per pixel (before the fp16 store), out_h: 2 halves per pixel.
int orc_motion_vectors(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const uint32_t* ids, float* out_f, uint16_t* out_h) {
	const vkv_MeshletDraw* draws = (const vkv_MeshletDraw*)pc->drawBuffer;
	const vkv_Primitive* prims = (const vkv_Primitive*)pc->primitiveBuffer;
	const vkv_Camera& cam = *(const vkv_Camera*)pc->cameraBuffer;
	const float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
	for (uint32_t y = 0; y < H; ++y)
		for (uint32_t x = 0; x < W; ++x) {
			const size_t i = (size_t)y * W + x;
			float mv[2] = {0.f, 0.f};                                                             // the clear value
			const uint32_t id = ids[i];
			if (id != VKV_VISBUFFER_CLEAR && pc->meshletDrawCount) {
				const uint32_t drawIndex = id >> VKV_TRIANGLE_BITS, tri = id & ((1u << VKV_TRIANGLE_BITS) - 1u);
				const vkv_MeshletDraw d = draws[drawIndex];
				const vkv_Primitive& prim = prims[d.primitiveIndex];
				const vkv_Meshlet& ml = ((const vkv_Meshlet*)prim.meshletBuffer)[d.meshletIndex];
				const uint8_t* t3 = (const uint8_t*)prim.primitiveIndexBuffer + ml.triangleOffset + tri * 3;
				const uint32_t* vidx = (const uint32_t*)prim.vertexIndexBuffer + ml.vertexOffset;
				const vkv_Vertex* verts = (const vkv_Vertex*)prim.vertexBuffer;
				const float* T = (const float*)pc->transformBuffer + (size_t)d.transformIndex * 16;
				float mvp[16], prevMvp[16];
				mul44m(cam.viewProjection, T, mvp);                                               // mesh.glsl:44
				mul44m(cam.prevViewProjection, T, prevMvp);                                       // mesh.glsl:45
				V4 pos[3], prev[3];
				for (int k = 0; k < 3; ++k) {
					const float* p = verts[vidx[t3[k]]].position;
					pos[k] = mul44(mvp, V4{p[0], p[1], p[2], 1.0f});                              // mesh.glsl:61-62
					prev[k] = mul44(prevMvp, V4{p[0], p[1], p[2], 1.0f});                         // mesh.glsl:63
				}
				const float nx = (((float)x + 0.5f) - hw) / hw, ny = (((float)y + 0.5f) - hh) / hh; // the pixel centre in NDC
				float l[3];
				for (int k = 0; k < 3; ++k) {
					const V4 &u = pos[(k + 1) % 3], &v = pos[(k + 2) % 3];
					l[k] = ((u.y * v.w - u.w * v.y) * nx + (u.w * v.x - u.x * v.w) * ny) + (u.x * v.y - u.y * v.x);
				}
				const float Px = (l[0] * pos[0].x + l[1] * pos[1].x) + l[2] * pos[2].x, Py = (l[0] * pos[0].y + l[1] * pos[1].y) + l[2] * pos[2].y;
				const float Pw = (l[0] * pos[0].w + l[1] * pos[1].w) + l[2] * pos[2].w;
				const float Qx = (l[0] * prev[0].x + l[1] * prev[1].x) + l[2] * prev[2].x, Qy = (l[0] * prev[0].y + l[1] * prev[1].y) + l[2] * prev[2].y;
				const float Qw = (l[0] * prev[0].w + l[1] * prev[1].w) + l[2] * prev[2].w;
				mv[0] = ((Qx / Qw) * 0.5f - 0.5f) - ((Px / Pw) * 0.5f - 0.5f);                    // frag.glsl:38
				mv[1] = ((Qy / Qw) * 0.5f - 0.5f) - ((Py / Pw) * 0.5f - 0.5f);
			}
			if (out_f) { out_f[i * 2] = mv[0]; out_f[i * 2 + 1] = mv[1]; }
			if (out_h) { out_h[i * 2] = to_half(mv[0]); out_h[i * 2 + 1] = to_half(mv[1]); }
		}
	return 0;
}

int orc_hiz(uint32_t W, uint32_t H, const float* depth, float* pyramid, int threads) {
	Pyr pyr;
	pyr.levels = orc_pyramid_layout(W, H, pyr.off, pyr.w, pyr.h, &pyr.total);
	// application.cpp:964-979: view 0 = depth image, view i = pyramid mip i-1; dispatch i writes levelSize = res >> i
	for (uint32_t i = 1; i <= pyr.levels; ++i) {
		const uint32_t dw = W >> i, dh = H >> i;
		if (dw == 0 || dh == 0) continue; // zero-sized dispatch: the mip keeps its previous contents
		const float* src = (i == 1) ? depth : pyramid + pyr.off[i - 2];
		const uint32_t sw = (i == 1) ? W : pyr.w[i - 2], sh = (i == 1) ? H : pyr.h[i - 2];
		float* dst = pyramid + pyr.off[i - 1];
		const uint32_t dstride = pyr.w[i - 1];
		parallel_for(dh, (dw * dh > 4096) ? threads : 1, [&](size_t b, size_t e, int) {
			for (size_t y = b; y < e; ++y)
				for (uint32_t x = 0; x < dw; ++x) {
					// hiz_reduce.comp.glsl:28 : texture(src, (vec2(pos) + 0.5) / imageSize)
					float u = ((float)x + 0.5f) / (float)dw;
					float v = ((float)y + 0.5f) / (float)dh;
					dst[y * dstride + x] = sample_min(src, sw, sh, u, v, nullptr);
				}
		});
	}
	return 0;
}

} // extern "C"
