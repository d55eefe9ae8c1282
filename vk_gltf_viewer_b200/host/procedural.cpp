// procedural.cpp — deterministic synthetic scenes for the BASELINE.json configs.
// There is no network for glTF sample assets, and fastgltf's parser cannot be built here (simdjson missing),
// so the scenes are generated in memory and fed through the same host pipeline a parsed glTF would take
// (vkvh_scene_add_primitive / add_node_trs / finalize == assets.cpp:288-373 + world.cpp:187-345).
#include "scene.hpp"

#include <atomic>
#include <thread>

#include <algorithm>
#include <cmath>
#include <functional>
#include <unordered_map>

using namespace vkvh;

namespace {

constexpr float kPi = 3.14159265358979323846f;

struct Mesh {
	std::vector<float> pos;
	std::vector<uint32_t> idx;
	uint32_t add(float x, float y, float z) {
		pos.push_back(x); pos.push_back(y); pos.push_back(z);
		return (uint32_t)(pos.size() / 3 - 1);
	}
	void tri(uint32_t a, uint32_t b, uint32_t c, bool flip) {
		idx.push_back(a);
		if (flip) { idx.push_back(c); idx.push_back(b); } else { idx.push_back(b); idx.push_back(c); }
	}
};

// Lofted grid: (nu+1)x(nv+1) vertices from f(u,v), u,v in [0,1]; 2*nu*nv triangles.  Normal = dP/du x dP/dv (flip reverses).
void loft(Mesh& m, uint32_t nu, uint32_t nv, const std::function<void(float, float, float*)>& f, bool flip) {
	const uint32_t base = (uint32_t)(m.pos.size() / 3);
	for (uint32_t j = 0; j <= nv; ++j)
		for (uint32_t i = 0; i <= nu; ++i) {
			float p[3];
			f((float)i / (float)nu, (float)j / (float)nv, p);
			m.add(p[0], p[1], p[2]);
		}
	for (uint32_t j = 0; j < nv; ++j)
		for (uint32_t i = 0; i < nu; ++i) {
			uint32_t a = base + j * (nu + 1) + i, b = a + 1, c = a + nu + 1, d = c + 1;
			m.tri(a, b, c, flip);
			m.tri(b, d, c, flip);
		}
}

// cheap deterministic value noise in [0,1)
float hash2(int x, int y, uint32_t seed) {
	uint32_t h = (uint32_t)x * 0x8da6b343u ^ (uint32_t)y * 0xd8163841u ^ seed * 0xcb1ab31fu;
	h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
	return (float)(h >> 8) * (1.0f / 16777216.0f);
}
float vnoise(float x, float y, uint32_t seed) {
	int xi = (int)std::floor(x), yi = (int)std::floor(y);
	float fx = x - (float)xi, fy = y - (float)yi;
	float sx = fx * fx * (3.0f - 2.0f * fx), sy = fy * fy * (3.0f - 2.0f * fy);
	float a = hash2(xi, yi, seed), b = hash2(xi + 1, yi, seed), c = hash2(xi, yi + 1, seed), d = hash2(xi + 1, yi + 1, seed);
	return (a + (b - a) * sx) + ((c + (d - c) * sx) - (a + (b - a) * sx)) * sy;
}

void set_bounds(vkvh_scene* s) {
	float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
	for (size_t i = 0; i < s->draws.size(); ++i) {
		const auto& d = s->draws[i];
		const auto& ml = s->primitives[d.primitiveIndex].meshlets[d.meshletIndex];
		const float* T = &s->transforms[(size_t)d.transformIndex * 16];
		for (int c = 0; c < 8; ++c) {
			float p[3] = {ml.aabbCenter[0] + ((c & 1) ? 1.f : -1.f) * ml.aabbExtents[0], ml.aabbCenter[1] + ((c & 2) ? 1.f : -1.f) * ml.aabbExtents[1],
			              ml.aabbCenter[2] + ((c & 4) ? 1.f : -1.f) * ml.aabbExtents[2]};
			for (int k = 0; k < 3; ++k) {
				float w = T[k] * p[0] + T[4 + k] * p[1] + T[8 + k] * p[2] + T[12 + k];
				mn[k] = std::min(mn[k], w); mx[k] = std::max(mx[k], w);
			}
		}
	}
	for (int k = 0; k < 3; ++k) { s->boundsMin[k] = mn[k]; s->boundsMax[k] = mx[k]; }
}

} // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------------------
// cfg 1: class-I geodesic icosphere, 20*f^2 triangles, 10*f^2+2 welded vertices, unit radius, CCW outward.
// ------------------------------------------------------------------------------------------------------------
vkvh_scene* vkvh_scene_icosphere(uint32_t f) {
	if (f == 0) f = 1;
	const float t = (1.0f + std::sqrt(5.0f)) / 2.0f;
	const float V[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
	                        {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
	const int F[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
	                      {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
	Mesh m;
	std::unordered_map<uint64_t, uint32_t> weld;
	auto emit = [&](double x, double y, double z) {
		double l = std::sqrt(x * x + y * y + z * z);
		return m.add((float)(x / l), (float)(y / l), (float)(z / l));
	};
	auto point = [&](int face, uint32_t a, uint32_t b, uint32_t c) -> uint32_t { // weights of the face's corners, a+b+c=f
		const int A = F[face][0], B = F[face][1], C = F[face][2];
		uint64_t key;
		int c0, c1; uint32_t w0, w1; // canonical edge description
		int nz = (a != 0) + (b != 0) + (c != 0);
		if (nz == 1) {
			int v = a ? A : (b ? B : C);
			key = (1ull << 60) | (uint64_t)v;
			auto it = weld.find(key);
			if (it != weld.end()) return it->second;
			return weld[key] = emit(V[v][0], V[v][1], V[v][2]);
		}
		if (nz == 2) {
			if (c == 0) { c0 = A; w0 = a; c1 = B; w1 = b; } else if (b == 0) { c0 = A; w0 = a; c1 = C; w1 = c; } else { c0 = B; w0 = b; c1 = C; w1 = c; }
			if (c0 > c1) { std::swap(c0, c1); std::swap(w0, w1); }
			key = (2ull << 60) | ((uint64_t)c0 << 40) | ((uint64_t)c1 << 32) | w1;
			auto it = weld.find(key);
			if (it != weld.end()) return it->second;
			return weld[key] = emit((double)V[c0][0] * w0 + (double)V[c1][0] * w1, (double)V[c0][1] * w0 + (double)V[c1][1] * w1,
			                        (double)V[c0][2] * w0 + (double)V[c1][2] * w1);
		}
		key = (3ull << 60) | ((uint64_t)face << 48) | ((uint64_t)b << 24) | c;
		auto it = weld.find(key);
		if (it != weld.end()) return it->second;
		return weld[key] = emit((double)V[A][0] * a + (double)V[B][0] * b + (double)V[C][0] * c, (double)V[A][1] * a + (double)V[B][1] * b + (double)V[C][1] * c,
		                        (double)V[A][2] * a + (double)V[B][2] * b + (double)V[C][2] * c);
	};
	for (int face = 0; face < 20; ++face)
		for (uint32_t i = 0; i < f; ++i)
			for (uint32_t j = 0; j + i < f; ++j) {
				uint32_t p00 = point(face, f - i - j, i, j), p10 = point(face, f - i - j - 1, i + 1, j), p01 = point(face, f - i - j - 1, i, j + 1);
				m.tri(p00, p10, p01, false);
				if (i + j + 1 < f) {
					uint32_t p11 = point(face, f - i - j - 2, i + 1, j + 1);
					m.tri(p10, p11, p01, false);
				}
			}
	vkvh_scene* s = vkvh_scene_new();
	int32_t prim = vkvh_scene_add_primitive(s, m.pos.data(), (uint32_t)(m.pos.size() / 3), m.idx.data(), (uint32_t)m.idx.size(), 0);
	vkvh_scene_add_node_trs(s, -1, prim, nullptr, nullptr, nullptr);
	vkvh_scene_finalize(s);
	s->kind = 1;
	set_bounds(s);
	return s;
}

// ------------------------------------------------------------------------------------------------------------
// cfg 2: procedural atrium, KHR_mesh_quantization-style int16 positions + node scale/offset.
// detail d: floor 2d^2 + vault 4d^2 + 2 walls d^2 each + 16 columns d^2/4 each + 8 arches d^2/2 each = 16 d^2 triangles
// (d = 128 -> 262,144).  28 mesh nodes, 5 unique primitives.
// ------------------------------------------------------------------------------------------------------------
vkvh_scene* vkvh_scene_atrium(uint32_t d) {
	if (d < 8) d = 8;
	vkvh_scene* s = vkvh_scene_new();
	const float white[4] = {0.8f, 0.8f, 0.75f, 1.0f}, red[4] = {0.7f, 0.3f, 0.2f, 1.0f};
	const uint32_t matStone = vkvh_scene_add_material(s, white, 0);
	const uint32_t matArch = vkvh_scene_add_material(s, red, 1); // double sided: exercises mesh.glsl:99-102
	// Quantise a mesh authored in the unit box [-1,1]^3 to int16 (non-normalised accessor).
	auto add_q = [&](const Mesh& m, uint32_t material) {
		std::vector<int16_t> q(m.pos.size());
		for (size_t i = 0; i < m.pos.size(); ++i) q[i] = (int16_t)std::lrintf(std::min(1.0f, std::max(-1.0f, m.pos[i])) * 32767.0f);
		return vkvh_scene_add_primitive_i16(s, q.data(), (uint32_t)(q.size() / 3), 0, m.idx.data(), (uint32_t)m.idx.size(), material);
	};
	const float inv = 1.0f / 32767.0f;
	auto node = [&](int32_t prim, float tx, float ty, float tz, float sx, float sy, float sz, float yaw) {
		float t[3] = {tx, ty, tz}, r[4] = {0, std::sin(yaw * 0.5f), 0, std::cos(yaw * 0.5f)}, sc[3] = {sx * inv, sy * inv, sz * inv};
		vkvh_scene_add_node_trs(s, -1, prim, t, r, sc);
	};
	// floor: y = gentle tiles, normal +y
	Mesh floorM;
	loft(floorM, d, d, [&](float u, float v, float* p) {
		p[0] = u * 2 - 1; p[2] = v * 2 - 1;
		p[1] = 0.02f * std::sin(u * 40.0f) * std::sin(v * 120.0f);
	}, true);
	int32_t pFloor = add_q(floorM, matStone);
	node(pFloor, 0, 0, 0, 10, 1, 30, 0);
	// vault: half cylinder, normal pointing inwards (down)
	Mesh vaultM;
	loft(vaultM, 2 * d, d, [&](float u, float v, float* p) {
		float a = u * kPi;
		p[0] = std::cos(a); p[1] = std::sin(a); p[2] = v * 2 - 1;
	}, true);
	int32_t pVault = add_q(vaultM, matStone);
	node(pVault, 0, 8, 0, 10, 4, 30, 0);
	// wall: plane x = -1 facing +x; second instance rotated 180 deg about y
	Mesh wallM;
	loft(wallM, d, d / 2, [&](float u, float v, float* p) {
		p[0] = -1.0f + 0.01f * std::sin(u * 60.0f); p[1] = v * 2 - 1; p[2] = u * 2 - 1;
	}, true);
	int32_t pWall = add_q(wallM, matStone);
	node(pWall, -0.0f, 4, 0, 10, 4, 30, 0);
	node(pWall, 0.0f, 4, 0, 10, 4, 30, kPi);
	// column: cylinder, outward normals
	Mesh colM;
	loft(colM, d / 2, d / 4, [&](float u, float v, float* p) {
		float a = u * 2 * kPi;
		float r = 0.8f + 0.2f * std::cos(a * 12.0f) * 0.3f; // fluting
		p[0] = r * std::cos(a); p[2] = r * std::sin(a); p[1] = v * 2 - 1;
	}, true);
	int32_t pCol = add_q(colM, matStone);
	for (int i = 0; i < 8; ++i) {
		float z = -26.0f + (float)i * (52.0f / 7.0f);
		node(pCol, -6, 4, z, 0.6f, 4, 0.6f, 0);
		node(pCol, 6, 4, z, 0.6f, 4, 0.6f, 0);
	}
	// arch: half torus across the nave, double sided
	Mesh archM;
	loft(archM, d / 2, d / 2, [&](float u, float v, float* p) {
		float a = u * kPi, b = v * 2 * kPi;
		float R = 0.9f, r = 0.1f;
		p[0] = (R + r * std::cos(b)) * std::cos(a); p[1] = (R + r * std::cos(b)) * std::sin(a); p[2] = r * std::sin(b);
	}, false);
	int32_t pArch = add_q(archM, matArch);
	for (int i = 0; i < 8; ++i) {
		float z = -26.0f + (float)i * (52.0f / 7.0f);
		node(pArch, 0, 8, z, 6.6f, 3.0f, 5.0f, 0);
	}
	vkvh_scene_finalize(s);
	s->kind = 2;
	set_bounds(s);
	return s;
}

// ------------------------------------------------------------------------------------------------------------
// cfg 3 / 5: nx x ny x nz lattice of ONE q x q-quad heightfield patch (2 q^2 triangles per instance).
// q = 224, 10x10x10 -> 100,352,000 triangles in 1000 instances; 22x22x21 -> 1.02 B.
// ------------------------------------------------------------------------------------------------------------
vkvh_scene* vkvh_scene_lattice(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t q, uint64_t seed) {
	vkvh_scene* s = vkvh_scene_new();
	Mesh m;
	const uint32_t sd = (uint32_t)(seed ^ (seed >> 32));
	loft(m, q, q, [&](float u, float v, float* p) {
		p[0] = u - 0.5f; p[2] = v - 0.5f;
		p[1] = 0.06f * vnoise(u * 9.0f, v * 9.0f, sd) + 0.015f * vnoise(u * 37.0f, v * 37.0f, sd + 1);
	}, true);
	int32_t prim = vkvh_scene_add_primitive(s, m.pos.data(), (uint32_t)(m.pos.size() / 3), m.idx.data(), (uint32_t)m.idx.size(), 0);
	for (uint32_t iy = 0; iy < ny; ++iy)
		for (uint32_t iz = 0; iz < nz; ++iz)
			for (uint32_t ix = 0; ix < nx; ++ix) {
				float t[3] = {(float)ix, (float)iy * 0.35f, (float)iz};
				vkvh_scene_add_node_trs(s, -1, prim, t, nullptr, nullptr);
			}
	vkvh_scene_finalize(s);
	s->kind = 3;
	set_bounds(s);
	return s;
}

// ------------------------------------------------------------------------------------------------------------
// cfg 4: city of nbx x nby UNIQUE buildings (no instancing) + tessellated ground.
// ------------------------------------------------------------------------------------------------------------
constexpr float kBuildingQuant = 512.0f, kGroundQuant = 64.0f; // int16 units per scene unit (buildings reach 40, the ground 300)
static vkvh_scene* city_impl(uint32_t nbx, uint32_t nby, uint32_t target, uint64_t seed, bool quantized);
vkvh_scene* vkvh_scene_city(uint32_t nbx, uint32_t nby, uint32_t target, uint64_t seed) { return city_impl(nbx, nby, target, seed, false); }
vkvh_scene* vkvh_scene_city_quantized(uint32_t nbx, uint32_t nby, uint32_t target, uint64_t seed) { return city_impl(nbx, nby, target, seed, true); }

static vkvh_scene* city_impl(uint32_t nbx, uint32_t nby, uint32_t target, uint64_t seed, bool quantized) {
	vkvh_scene* s = vkvh_scene_new();
	SplitMix64 rng(seed);
	// facade n x m quads on 4 sides + n x n roof: 8nm + 2n^2 ~= target, with m ~= 2.25 n
	uint32_t n = (uint32_t)std::max(2.0, std::floor(std::sqrt((double)target / 20.0)));
	uint32_t mq = (uint32_t)std::max(2.0, std::floor(((double)target - 2.0 * n * n) / (8.0 * n)));
	const float lot = 10.0f;
	// geometry first (sequential: one seeded random stream), then the meshlet builds of all buildings side by side, then the
	// primitives and nodes in the original order
	std::vector<Mesh> meshes;
	struct Placement { float t[3]; float r[4]; };
	std::vector<Placement> places;
	for (uint32_t by = 0; by < nby; ++by)
		for (uint32_t bx = 0; bx < nbx; ++bx) {
			const float w = rng.range(2.5f, 4.2f), dpt = rng.range(2.5f, 4.2f), h = rng.range(6.0f, 40.0f);
			const float relief = rng.range(0.02f, 0.12f);
			const float fu = std::floor(rng.range(4.0f, 12.0f)), fv = std::floor(rng.range(8.0f, 30.0f));
			const uint32_t bs = (uint32_t)rng.next();
			Mesh m;
			auto bump = [&](float u, float v) {
				float a = std::sin(u * fu * kPi), b = std::sin(v * fv * kPi);
				return relief * (a * a * b * b) + 0.02f * hash2((int)(u * 64), (int)(v * 64), bs);
			};
			// four facades (outward normals), each its own lofted grid
			loft(m, n, mq, [&](float u, float v, float* p) { p[0] = (u * 2 - 1) * w; p[1] = v * h; p[2] = dpt + bump(u, v); }, false);
			loft(m, n, mq, [&](float u, float v, float* p) { p[0] = (1 - u * 2) * w; p[1] = v * h; p[2] = -dpt - bump(u, v); }, false);
			loft(m, n, mq, [&](float u, float v, float* p) { p[2] = (1 - u * 2) * dpt; p[1] = v * h; p[0] = w + bump(u, v); }, false);
			loft(m, n, mq, [&](float u, float v, float* p) { p[2] = (u * 2 - 1) * dpt; p[1] = v * h; p[0] = -w - bump(u, v); }, false);
			loft(m, n, n, [&](float u, float v, float* p) { p[0] = (u * 2 - 1) * w; p[2] = (v * 2 - 1) * dpt; p[1] = h + 0.3f * bump(u, v); }, true);
			Placement pl;
			pl.t[0] = ((float)bx - (float)(nbx - 1) * 0.5f) * lot; pl.t[1] = 0.0f; pl.t[2] = ((float)by - (float)(nby - 1) * 0.5f) * lot;
			float yaw = rng.range(-0.2f, 0.2f);
			pl.r[0] = 0; pl.r[1] = std::sin(yaw * 0.5f); pl.r[2] = 0; pl.r[3] = std::cos(yaw * 0.5f);
			meshes.push_back(std::move(m));
			places.push_back(pl);
		}
	{
		Mesh g;
		const float ex = (float)nbx * lot * 0.6f, ez = (float)nby * lot * 0.6f;
		loft(g, 256, 256, [&](float u, float v, float* p) { p[0] = (u * 2 - 1) * ex; p[2] = (v * 2 - 1) * ez; p[1] = 0.0f; }, true);
		meshes.push_back(std::move(g));
	}
	std::vector<vkvh::PrimitiveData> built(meshes.size());
	std::vector<char> ok(meshes.size(), 0);
	{
		std::atomic<size_t> next{0};
		auto worker = [&]() {
			for (size_t i; (i = next.fetch_add(1)) < meshes.size();) {
				const Mesh& m = meshes[i];
				const uint32_t nv = (uint32_t)(m.pos.size() / 3);
				if (!quantized) {
					ok[i] = vkvh::build_primitive(built[i], vkvh::vertices_from_positions(m.pos.data(), nv), m.idx.data(), (uint32_t)m.idx.size(), 0) ? 1 : 0;
					continue;
				}
				// KHR_mesh_quantization as gltfpack writes it: SHORT positions in units of 1 / qscale, the node scale undoes it
				const float qs = i < places.size() ? kBuildingQuant : kGroundQuant;
				std::vector<int16_t> q((size_t)nv * 4, 0);
				std::vector<float> fq(m.pos.size());
				for (uint32_t v = 0; v < nv; ++v)
					for (int k = 0; k < 3; ++k) {
						const long r = std::lrintf(m.pos[v * 3 + k] * qs);
						q[v * 4 + k] = (int16_t)std::max(-32767L, std::min(32767L, r));
						fq[v * 3 + k] = (float)q[v * 4 + k]; // convertComponent<float, int16_t>, not normalized
					}
				ok[i] = vkvh::build_primitive(built[i], vkvh::vertices_from_positions(fq.data(), nv), m.idx.data(), (uint32_t)m.idx.size(), 0) ? 1 : 0;
				built[i].qpos = std::move(q);
				built[i].qnormalized = false;
			}
		};
		const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
		std::vector<std::thread> pool;
		for (unsigned t = 1; t < nt; ++t) pool.emplace_back(worker);
		worker();
		for (auto& th : pool) th.join();
	}
	for (size_t i = 0; i < meshes.size(); ++i) {
		if (!ok[i]) { vkvh_scene_free(s); return nullptr; }
		const int32_t prim = vkvh::add_built_primitive(s, std::move(built[i]));
		const float inv = 1.0f / (i < places.size() ? kBuildingQuant : kGroundQuant);
		const float sc[3] = {inv, inv, inv};
		if (i < places.size()) vkvh_scene_add_node_trs(s, -1, prim, places[i].t, places[i].r, quantized ? sc : nullptr);
		else vkvh_scene_add_node_trs(s, -1, prim, nullptr, nullptr, quantized ? sc : nullptr);
	}
	vkvh_scene_finalize(s);
	s->kind = 4;
	set_bounds(s);
	return s;
}

void vkvh_scene_default_view(const vkvh_scene* s, uint32_t view, uint32_t nviews, float eye[3], float center[3]) {
	float c[3], e[3];
	for (int k = 0; k < 3; ++k) { c[k] = (s->boundsMin[k] + s->boundsMax[k]) * 0.5f; e[k] = (s->boundsMax[k] - s->boundsMin[k]) * 0.5f; }
	if (nviews == 0) nviews = 1;
	const float ang = 2.0f * kPi * (float)view / (float)nviews;
	switch (s->kind) {
		case 1: // icosphere: eye on +Z at distance 3 (view 0), orbiting for other views
			eye[0] = 3.0f * std::sin(ang); eye[1] = 0.0f; eye[2] = 3.0f * std::cos(ang);
			center[0] = center[1] = center[2] = 0.0f;
			break;
		case 2: // atrium: interior view down the nave
			eye[0] = 0.5f + 2.0f * std::sin(ang); eye[1] = 2.5f; eye[2] = 27.0f;
			center[0] = 0.0f; center[1] = 4.0f; center[2] = 0.0f;
			break;
		case 3: { // lattice: outside a corner, looking along the diagonal
			float r = 1.0f + 0.15f * std::sin(ang);
			eye[0] = s->boundsMax[0] + 0.6f * r; eye[1] = s->boundsMax[1] + 0.25f * e[0] * r; eye[2] = s->boundsMax[2] + 0.6f * r;
			center[0] = c[0] + 0.35f * e[0]; center[1] = c[1]; center[2] = c[2] + 0.35f * e[2];
			break;
		}
		case 4: { // city: circle of radius 1.5 x extent at height 0.3 x extent
			float ext = std::max(e[0], e[2]);
			eye[0] = c[0] + 1.5f * ext * std::cos(ang); eye[2] = c[2] + 1.5f * ext * std::sin(ang); eye[1] = 0.3f * ext;
			center[0] = c[0]; center[1] = 0.0f; center[2] = c[2];
			break;
		}
		default: {
			float r = 2.0f * std::max(e[0], std::max(e[1], e[2])) + 1.0f;
			eye[0] = c[0] + r * std::sin(ang); eye[1] = c[1] + 0.5f * r; eye[2] = c[2] + r * std::cos(ang);
			center[0] = c[0]; center[1] = c[1]; center[2] = c[2];
		}
	}
}

} // extern "C"
