// meshlet_builder.cpp — the round-1 meshlet packer, kept as an alternative workload (vkvh_select_builder(1)).
//
// The default builder is host/clusterizer.cpp, which reproduces the partition the reference uploads
// (meshopt_buildMeshlets(…, 64, 124, 0.0f) + meshopt_optimizeMeshlet, assets.cpp:322-346) byte for byte.  This one orders the
// triangles along a Morton curve of their centroids and packs them greedily until the 64-vertex / 124-triangle limit is hit:
// on regular grids it yields vertex-limited 64 v / ~67 t meshlets (40 % more MeshletDraws than the reference's ~95 t ones),
// i.e. a harder input for the same image — useful for A/B measurements, valid for the hot path (any partition within the
// limits is).  Output layout is the one the reference uploads: u32 global vertex index per local slot, 3 u8 local slots per
// triangle, each meshlet's triangle bytes starting on a 4-byte boundary (assets.cpp:339).
#include "scene.hpp"

#include <algorithm>
#include <cfloat>

namespace vkvh {

namespace {
inline uint32_t part1by2(uint32_t x) {
	x &= 0x3ff;
	x = (x | (x << 16)) & 0x030000FF;
	x = (x | (x << 8)) & 0x0300F00F;
	x = (x | (x << 4)) & 0x030C30C3;
	x = (x | (x << 2)) & 0x09249249;
	return x;
}
} // namespace

void build_meshlets_builtin(const std::vector<vkv_Vertex>& vertices, const std::vector<uint32_t>& indices,
                            std::vector<MeshletRec>& meshlets, std::vector<uint32_t>& meshletVertices,
                            std::vector<uint8_t>& meshletTriangles) {
	const size_t ntri = indices.size() / 3;
	const size_t nv = vertices.size();
	// centroid Morton keys over the primitive's AABB
	float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	for (const auto& v : vertices)
		for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], v.position[k]); mx[k] = std::max(mx[k], v.position[k]); }
	float inv[3];
	for (int k = 0; k < 3; ++k) inv[k] = (mx[k] > mn[k]) ? 1023.0f / (mx[k] - mn[k]) : 0.0f;
	std::vector<std::pair<uint32_t, uint32_t>> order(ntri);
	for (size_t t = 0; t < ntri; ++t) {
		uint32_t q[3];
		for (int k = 0; k < 3; ++k) {
			float c = (vertices[indices[t * 3]].position[k] + vertices[indices[t * 3 + 1]].position[k] + vertices[indices[t * 3 + 2]].position[k]) * (1.0f / 3.0f);
			float f = (c - mn[k]) * inv[k];
			q[k] = (uint32_t)std::min(1023.0f, std::max(0.0f, f));
		}
		order[t] = {part1by2(q[0]) | (part1by2(q[1]) << 1) | (part1by2(q[2]) << 2), (uint32_t)t};
	}
	std::sort(order.begin(), order.end());

	std::vector<int16_t> slot(nv, -1);
	MeshletRec cur{};
	cur.vertex_offset = 0; cur.triangle_offset = 0; cur.vertex_count = 0; cur.triangle_count = 0;
	auto flush = [&]() {
		if (cur.triangle_count == 0) return;
		for (uint32_t i = 0; i < cur.vertex_count; ++i) slot[meshletVertices[cur.vertex_offset + i]] = -1;
		while (meshletTriangles.size() & 3) meshletTriangles.push_back(0);
		meshlets.push_back(cur);
		cur.vertex_offset = (uint32_t)meshletVertices.size();
		cur.triangle_offset = (uint32_t)meshletTriangles.size();
		cur.vertex_count = cur.triangle_count = 0;
	};
	for (size_t o = 0; o < ntri; ++o) {
		const uint32_t t = order[o].second;
		const uint32_t a = indices[t * 3], b = indices[t * 3 + 1], c = indices[t * 3 + 2];
		uint32_t fresh = (slot[a] < 0) + (slot[b] < 0 && b != a) + (slot[c] < 0 && c != a && c != b);
		if (cur.vertex_count + fresh > VKV_MAX_VERTICES || cur.triangle_count + 1 > VKV_MAX_MESHLET_TRIANGLES) flush();
		for (uint32_t v : {a, b, c}) {
			if (slot[v] < 0) {
				slot[v] = (int16_t)cur.vertex_count++;
				meshletVertices.push_back(v);
			}
			meshletTriangles.push_back((uint8_t)slot[v]);
		}
		cur.triangle_count++;
	}
	flush();
}

} // namespace vkvh
