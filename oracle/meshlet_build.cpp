/* meshlet_build.cpp — CPU restatement of the meshlet partition and bounds the device-side builder produces.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Partition: meshoptimizer's SCAN builder, submodules/meshoptimizer/src/clusterizer.cpp:224-298 (finishMeshlet, appendMeshlet)
 * and :673-708 (meshopt_buildMeshletsScan): triangles are taken in index-buffer order; the current meshlet is closed when
 * the next triangle's not-yet-used corners (counted per corner, so a repeated corner counts twice) would exceed max_vertices
 * or when it already holds max_triangles; local vertex numbers are first-appearance order; each meshlet's triangle bytes are
 * padded with zeros to a multiple of 4.  PARITY PINNED for this function: tests/test_meshlet_build.py checks it byte for byte
 * against meshopt_buildMeshletsScan of the reference's meshoptimizer built from source (oracle/_ref) and against the frozen
 * vectors in tests/golden/meshlet_scan.npz.
 * NOT what the reference calls: assets.cpp:331 uses meshopt_buildMeshlets (kd-tree + adjacency greedy, clusterizer.cpp:535-670)
 * followed by meshopt_optimizeMeshlet; neither has a golden output, both are order heuristics, and any valid
 * partition renders the same image.  The scan partition is the one member of that family that parallelises exactly.
 *
 * Bounds: assets.cpp:349-372 — component-wise glm::min / glm::max over the meshlet's vertices starting from the first,
 * center = (min + max) * 0.5f, extents = max - center.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <vector>

extern "C" {

struct orc_meshlet { uint32_t vertex_offset, triangle_offset, vertex_count, triangle_count; };

size_t orc_meshlets_scan(orc_meshlet* meshlets, uint32_t* meshlet_vertices, uint8_t* meshlet_triangles, const uint32_t* indices,
                         size_t index_count, size_t vertex_count, size_t max_vertices, size_t max_triangles) {
	std::vector<uint8_t> used(vertex_count, 0xff);
	orc_meshlet m = {0, 0, 0, 0};
	size_t n = 0;
	auto close = [&]() {
		size_t off = m.triangle_offset + m.triangle_count * 3;
		while (off & 3) meshlet_triangles[off++] = 0;
		meshlets[n++] = m;
	};
	for (size_t i = 0; i < index_count; i += 3) {
		const uint32_t v[3] = {indices[i], indices[i + 1], indices[i + 2]};
		const unsigned extra = (used[v[0]] == 0xff) + (used[v[1]] == 0xff) + (used[v[2]] == 0xff);
		if (m.vertex_count + extra > max_vertices || m.triangle_count >= max_triangles) {
			for (uint32_t j = 0; j < m.vertex_count; ++j) used[meshlet_vertices[m.vertex_offset + j]] = 0xff;
			close();
			m.vertex_offset += m.vertex_count;
			m.triangle_offset += (m.triangle_count * 3 + 3) & ~3u;
			m.vertex_count = m.triangle_count = 0;
		}
		for (int k = 0; k < 3; ++k) {
			if (used[v[k]] == 0xff) {
				used[v[k]] = (uint8_t)m.vertex_count;
				meshlet_vertices[m.vertex_offset + m.vertex_count++] = v[k];
			}
			meshlet_triangles[m.triangle_offset + m.triangle_count * 3 + k] = used[v[k]];
		}
		m.triangle_count++;
	}
	if (m.triangle_count) close();
	return n;
}

/* positions: float[3] at `stride` bytes; out: extents[3] then center[3] per meshlet */
void orc_meshlet_bounds(const orc_meshlet* meshlets, size_t count, const uint32_t* meshlet_vertices, const void* positions, size_t stride,
                        float* out_extents_center) {
	for (size_t i = 0; i < count; ++i) {
		const orc_meshlet& m = meshlets[i];
		float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
		for (uint32_t j = 0; j < m.vertex_count; ++j) {
			const float* p = (const float*)((const char*)positions + (size_t)meshlet_vertices[m.vertex_offset + j] * stride);
			for (int k = 0; k < 3; ++k) {
				if (j == 0) { mn[k] = mx[k] = p[k]; continue; }
				mn[k] = p[k] < mn[k] ? p[k] : mn[k]; /* glm::min(x, y) = y < x ? y : x */
				mx[k] = mx[k] < p[k] ? p[k] : mx[k]; /* glm::max(x, y) = x < y ? y : x */
			}
		}
		for (int k = 0; k < 3; ++k) {
			const float c = (mn[k] + mx[k]) * 0.5f;
			out_extents_center[i * 6 + k] = mx[k] - c;
			out_extents_center[i * 6 + 3 + k] = c;
		}
	}
}

} // extern "C"
