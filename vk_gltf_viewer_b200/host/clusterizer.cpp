// clusterizer.cpp — the meshlet partition the reference actually uploads (input generator; not part of the per-frame path).
//
// The reference builds its meshlets with meshoptimizer 0.20 (submodule pinned by the reference tree):
//   meshopt_buildMeshlets(…, maxVertices 64, maxTriangles 124, cone_weight 0.0f)       assets.cpp:331-335
//   meshopt_optimizeMeshlet(…) per meshlet                                             assets.cpp:345
// and (disabled at assets.cpp:323, but named by the north star's cone cull) meshopt_computeMeshletBounds.
// meshoptimizer is a third-party dependency of the reference; it is not linked here.  This file restates its published
// algorithms — submodules/meshoptimizer/src/clusterizer.cpp: adjacency + kd-tree greedy growth (:29-76, :154-510, :535-671),
// intra-meshlet reordering (:886-974), bounding sphere + normal cone (:78-152, :712-884) — as one self-contained builder whose
// output is BYTE-IDENTICAL to that library built from the reference tree (tests/test_host.py compares every record, vertex
// list, triangle byte and bounds field against oracle/_ref/libmeshopt_ref.so on several meshes).  Because the partition is an
// order heuristic, "identical" means: the same fp32 operations in the same order (this file is compiled with
// -ffp-contract=off; the library's x86-64 build has no FMA either), the same tie-breaking, the same traversal orders.
#include "scene.hpp"

#include <cfloat>
#include <cmath>
#include <cstring>

namespace vkvh {
namespace {

constexpr uint8_t kUnused = 0xff;           // "vertex is not in the open meshlet" (meshlets hold <= 255 vertices)
constexpr uint32_t kNone = ~0u;

struct TriInfo { float c[3]; float n[3]; }; // centroid and unit normal of one triangle (the library's `Cone`)

// One build job.  All state the greedy loop touches lives here, struct-of-arrays.
class GreedyBuilder {
public:
	GreedyBuilder(const unsigned* indices, size_t indexCount, const float* positions, size_t vertexCount, size_t strideBytes,
	              size_t maxVertices, size_t maxTriangles, float coneWeight)
	    : idx_(indices), faces_(indexCount / 3), pos_(positions), nverts_(vertexCount), strideF_(strideBytes / sizeof(float)),
	      maxV_(maxVertices), maxT_(maxTriangles), coneWeight_(coneWeight) {}

	size_t run(vkvh_meshopt_Meshlet* out, unsigned* outVertices, unsigned char* outTriangles);

private:
	// ---- vertex -> incident triangles (clusterizer.cpp:29-76) --------------------------------------------------------------
	void buildAdjacency() {
		adjCount_.assign(nverts_, 0);
		adjOffset_.assign(nverts_, 0);
		adjData_.assign(faces_ * 3, 0);
		for (size_t i = 0; i < faces_ * 3; ++i) adjCount_[idx_[i]]++;
		unsigned running = 0;
		for (size_t v = 0; v < nverts_; ++v) { adjOffset_[v] = running; running += adjCount_[v]; }
		std::vector<unsigned> fill(adjOffset_); // the library bumps offsets while filling and repairs them afterwards: same lists
		for (size_t f = 0; f < faces_; ++f)
			for (int k = 0; k < 3; ++k) adjData_[fill[idx_[f * 3 + k]]++] = (unsigned)f;
	}
	// swap-with-last removal of `face` from the incidence list of each of its corners (:641-659); a repeated corner is visited
	// once per occurrence, exactly as the library does
	void dropFromAdjacency(unsigned face) {
		for (int k = 0; k < 3; ++k) {
			const unsigned v = idx_[face * 3 + k];
			unsigned* list = adjData_.data() + adjOffset_[v];
			const size_t n = adjCount_[v];
			for (size_t i = 0; i < n; ++i)
				if (list[i] == face) { list[i] = list[n - 1]; adjCount_[v]--; break; }
		}
	}

	// ---- per-triangle centroid / unit normal, total area (:182-222) ---------------------------------------------------------
	float computeTriangles() {
		tri_.resize(faces_);
		float area_sum = 0;
		for (size_t f = 0; f < faces_; ++f) {
			const float* p0 = pos_ + strideF_ * idx_[f * 3];
			const float* p1 = pos_ + strideF_ * idx_[f * 3 + 1];
			const float* p2 = pos_ + strideF_ * idx_[f * 3 + 2];
			const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
			const float e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
			const float nx = e1[1] * e2[2] - e1[2] * e2[1];
			const float ny = e1[2] * e2[0] - e1[0] * e2[2];
			const float nz = e1[0] * e2[1] - e1[1] * e2[0];
			const float area = sqrtf(nx * nx + ny * ny + nz * nz);
			const float inv = (area == 0.f) ? 0.f : 1.f / area;
			TriInfo& t = tri_[f];
			for (int k = 0; k < 3; ++k) t.c[k] = (p0[k] + p1[k] + p2[k]) / 3.f;
			t.n[0] = nx * inv; t.n[1] = ny * inv; t.n[2] = nz * inv;
			area_sum += area;
		}
		return area_sum;
	}

	// ---- kd-tree over the triangle centroids (:373-510) --------------------------------------------------------------------
	// Node encoding as in the library: leaf = axis 3 with `children` extra points stored in the nodes right behind it; branch =
	// split value, left subtree at +1, right subtree at +1 + children.
	struct KdNode { union { float split; unsigned index; }; unsigned axis : 2; unsigned children : 30; };

	size_t kdLeaf(size_t at, const unsigned* pts, size_t n) {
		kd_[at].index = pts[0]; kd_[at].axis = 3; kd_[at].children = (unsigned)(n - 1);
		for (size_t i = 1; i < n; ++i) { kd_[at + i].index = pts[i]; kd_[at + i].axis = 3; kd_[at + i].children = ~0u >> 2; }
		return at + n;
	}
	size_t kdBuild(size_t at, unsigned* pts, size_t n) {
		constexpr size_t kLeaf = 8;
		if (n <= kLeaf) return kdLeaf(at, pts, n);
		// running mean / variance per axis (Welford), the reciprocal count carried exactly as the library carries it
		float mean[3] = {0, 0, 0}, var[3] = {0, 0, 0};
		float count = 1, rcount = 1;
		for (size_t i = 0; i < n; ++i, count += 1.f, rcount = 1.f / count) {
			const float* p = tri_[pts[i]].c;
			for (int k = 0; k < 3; ++k) {
				const float d = p[k] - mean[k];
				mean[k] += d * rcount;
				var[k] += d * (p[k] - mean[k]);
			}
		}
		const unsigned axis = (var[0] >= var[1] && var[0] >= var[2]) ? 0u : (var[1] >= var[2] ? 1u : 2u);
		const float pivot = mean[axis];
		// partition (< pivot first) with the library's unconditional swap: the resulting ORDER inside each half matters
		size_t mid = 0;
		for (size_t i = 0; i < n; ++i) {
			const float v = tri_[pts[i]].c[axis];
			const unsigned t = pts[mid]; pts[mid] = pts[i]; pts[i] = t;
			mid += v < pivot;
		}
		if (mid <= kLeaf / 2 || mid >= n - kLeaf / 2) return kdLeaf(at, pts, n); // degenerate split: one fat leaf
		kd_[at].split = pivot;
		kd_[at].axis = axis;
		const size_t right = kdBuild(at + 1, pts, mid);
		kd_[at].children = (unsigned)(right - at - 1);
		return kdBuild(right, pts + mid, n - mid);
	}
	void kdNearest(unsigned at, const float q[3], unsigned& best, float& limit) const {
		const KdNode& node = kd_[at];
		if (node.axis == 3) {
			for (unsigned i = 0; i <= node.children; ++i) {
				const unsigned f = kd_[at + i].index;
				if (emitted_[f]) continue;
				const float* p = tri_[f].c;
				const float d2 = (p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2]);
				const float d = sqrtf(d2);
				if (d < limit) { best = f; limit = d; }
			}
			return;
		}
		const float delta = q[node.axis] - node.split;
		const unsigned first = (delta <= 0) ? 0u : node.children;
		const unsigned second = first ^ node.children;
		kdNearest(at + 1 + first, q, best, limit);
		if (fabsf(delta) <= limit) kdNearest(at + 1 + second, q, best, limit);
	}

	// ---- candidate selection among the triangles touching the open meshlet (:286-363) ----------------------------------------
	// geometric = score by distance to the meshlet's running centroid / cone; otherwise topological (fewest live neighbours)
	unsigned pickNeighbour(bool geometric, const TriInfo& cone, float coneWeight, unsigned* extraOut) const {
		unsigned best = kNone, bestExtra = 5;
		float bestScore = FLT_MAX;
		for (size_t i = 0; i < cur_.vertex_count; ++i) {
			const unsigned v = outV_[cur_.vertex_offset + i];
			const unsigned* list = adjData_.data() + adjOffset_[v];
			const size_t n = adjCount_[v];
			for (size_t j = 0; j < n; ++j) {
				const unsigned f = list[j];
				const unsigned a = idx_[f * 3], b = idx_[f * 3 + 1], c = idx_[f * 3 + 2];
				unsigned extra = (slot_[a] == kUnused) + (slot_[b] == kUnused) + (slot_[c] == kUnused);
				if (extra != 0) { // dangling triangles (a corner with no other live triangle) are promoted
					if (live_[a] == 1 || live_[b] == 1 || live_[c] == 1) extra = 0;
					extra++;
				}
				if (extra > bestExtra) continue;
				float score;
				if (geometric) {
					const TriInfo& t = tri_[f];
					const float d2 = (t.c[0] - cone.c[0]) * (t.c[0] - cone.c[0]) + (t.c[1] - cone.c[1]) * (t.c[1] - cone.c[1]) +
					                 (t.c[2] - cone.c[2]) * (t.c[2] - cone.c[2]);
					const float spread = t.n[0] * cone.n[0] + t.n[1] * cone.n[1] + t.n[2] * cone.n[2];
					const float k = 1.f - spread * coneWeight;
					const float kc = k < 1e-3f ? 1e-3f : k;
					score = (1 + sqrtf(d2) / expectedRadius_ * (1 - coneWeight)) * kc;
				} else {
					score = float(live_[a] + live_[b] + live_[c] - 3);
				}
				if (extra < bestExtra || score < bestScore) { best = f; bestExtra = extra; bestScore = score; }
			}
		}
		if (extraOut) *extraOut = bestExtra;
		return best;
	}

	// ---- append one triangle, closing the open meshlet first when it does not fit (:224-284) ---------------------------------
	bool append(unsigned a, unsigned b, unsigned c) {
		bool closed = false;
		const unsigned fresh = (slot_[a] == kUnused) + (slot_[b] == kUnused) + (slot_[c] == kUnused);
		if (cur_.vertex_count + fresh > maxV_ || cur_.triangle_count >= maxT_) {
			outM_[nOut_] = cur_;
			for (size_t j = 0; j < cur_.vertex_count; ++j) slot_[outV_[cur_.vertex_offset + j]] = kUnused;
			padTriangles();
			cur_.vertex_offset += cur_.vertex_count;
			cur_.triangle_offset += (cur_.triangle_count * 3 + 3) & ~3u;
			cur_.vertex_count = 0;
			cur_.triangle_count = 0;
			closed = true;
		}
		const unsigned corners[3] = {a, b, c};
		for (unsigned v : corners)
			if (slot_[v] == kUnused) {
				slot_[v] = (uint8_t)cur_.vertex_count;
				outV_[cur_.vertex_offset + cur_.vertex_count++] = v;
			}
		unsigned char* t = outT_ + cur_.triangle_offset + cur_.triangle_count * 3;
		t[0] = slot_[a]; t[1] = slot_[b]; t[2] = slot_[c];
		cur_.triangle_count++;
		return closed;
	}
	void padTriangles() { // zero bytes up to the next 4-byte boundary (assets.cpp:339 relies on it)
		for (size_t o = cur_.triangle_offset + cur_.triangle_count * 3; o & 3; ++o) outT_[o] = 0;
	}

	const unsigned* idx_; size_t faces_;
	const float* pos_; size_t nverts_, strideF_;
	size_t maxV_, maxT_;
	float coneWeight_, expectedRadius_ = 0;
	std::vector<unsigned> adjCount_, adjOffset_, adjData_, live_;
	std::vector<TriInfo> tri_;
	std::vector<KdNode> kd_;
	std::vector<uint8_t> emitted_, slot_;
	vkvh_meshopt_Meshlet cur_{};
	vkvh_meshopt_Meshlet* outM_ = nullptr; unsigned* outV_ = nullptr; unsigned char* outT_ = nullptr;
	size_t nOut_ = 0;
};

size_t GreedyBuilder::run(vkvh_meshopt_Meshlet* out, unsigned* outVertices, unsigned char* outTriangles) {
	outM_ = out; outV_ = outVertices; outT_ = outTriangles;
	buildAdjacency();
	live_ = adjCount_;
	emitted_.assign(faces_, 0);
	const float area = computeTriangles();
	// "each meshlet is a square patch": expected radius from the average triangle area (:563-565)
	const float avgArea = faces_ == 0 ? 0.f : area / float(faces_) * 0.5f;
	expectedRadius_ = sqrtf(avgArea * maxT_) * 0.5f;
	kd_.resize(faces_ * 2);
	if (faces_) {
		std::vector<unsigned> order(faces_);
		for (size_t i = 0; i < faces_; ++i) order[i] = (unsigned)i;
		kdBuild(0, order.data(), faces_);
	}
	slot_.assign(nverts_, kUnused);

	float acc[6] = {0, 0, 0, 0, 0, 0}; // sums of centroids and normals of the open meshlet
	for (;;) {
		// the open meshlet's centroid and (normalised) mean normal (:162-180)
		TriInfo cone;
		{
			const float cs = cur_.triangle_count == 0 ? 0.f : 1.f / float(cur_.triangle_count);
			for (int k = 0; k < 3; ++k) cone.c[k] = acc[k] * cs;
			float n[3] = {acc[3], acc[4], acc[5]};
			const float len2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
			const float ns = len2 == 0.f ? 0.f : 1.f / sqrtf(len2);
			for (int k = 0; k < 3; ++k) cone.n[k] = n[k] * ns;
		}
		unsigned extra = 0;
		unsigned next = pickNeighbour(true, cone, coneWeight_, &extra);
		// the best geometric candidate would overflow the meshlet: re-select by topology (:600-603)
		if (next != kNone && (cur_.vertex_count + extra > maxV_ || cur_.triangle_count >= maxT_)) next = pickNeighbour(false, cone, 0.f, nullptr);
		if (next == kNone && faces_) { // no live neighbour left: nearest unemitted triangle anywhere (:606-615)
			float limit = FLT_MAX;
			kdNearest(0, cone.c, next, limit);
		}
		if (next == kNone) break;
		const unsigned a = idx_[next * 3], b = idx_[next * 3 + 1], c = idx_[next * 3 + 2];
		if (append(a, b, c)) { nOut_++; std::memset(acc, 0, sizeof(acc)); }
		live_[a]--; live_[b]--; live_[c]--;
		dropFromAdjacency(next);
		for (int k = 0; k < 3; ++k) { acc[k] += tri_[next].c[k]; acc[3 + k] += tri_[next].n[k]; }
		emitted_[next] = 1;
	}
	if (cur_.triangle_count) { padTriangles(); outM_[nOut_++] = cur_; }
	return nOut_;
}

// ---- Ritter-style bounding sphere (:78-152) ----------------------------------------------------------------------------------
void boundingSphere(float out[4], const float (*pts)[3], size_t n) {
	size_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
	for (size_t i = 0; i < n; ++i)
		for (int k = 0; k < 3; ++k) {
			lo[k] = (pts[i][k] < pts[lo[k]][k]) ? i : lo[k];
			hi[k] = (pts[i][k] > pts[hi[k]][k]) ? i : hi[k];
		}
	float best2 = 0;
	int bestAxis = 0;
	for (int k = 0; k < 3; ++k) {
		const float* p = pts[lo[k]];
		const float* q = pts[hi[k]];
		const float d2 = (q[0] - p[0]) * (q[0] - p[0]) + (q[1] - p[1]) * (q[1] - p[1]) + (q[2] - p[2]) * (q[2] - p[2]);
		if (d2 > best2) { best2 = d2; bestAxis = k; }
	}
	const float* p = pts[lo[bestAxis]];
	const float* q = pts[hi[bestAxis]];
	float ctr[3] = {(p[0] + q[0]) / 2, (p[1] + q[1]) / 2, (p[2] + q[2]) / 2};
	float radius = sqrtf(best2) / 2;
	for (size_t i = 0; i < n; ++i) { // grow until every point fits
		const float* x = pts[i];
		const float d2 = (x[0] - ctr[0]) * (x[0] - ctr[0]) + (x[1] - ctr[1]) * (x[1] - ctr[1]) + (x[2] - ctr[2]) * (x[2] - ctr[2]);
		if (d2 > radius * radius) {
			const float d = sqrtf(d2);
			const float k = 0.5f + (radius / d) / 2;
			for (int c = 0; c < 3; ++c) ctr[c] = ctr[c] * k + x[c] * (1 - k);
			radius = (radius + d) / 2;
		}
	}
	out[0] = ctr[0]; out[1] = ctr[1]; out[2] = ctr[2]; out[3] = radius;
}

int quantizeSnorm8(float v) { // meshopt_quantizeSnorm(v, 8) (meshoptimizer.h:969-980)
	const float scale = 127.f;
	const float round = (v >= 0 ? 0.5f : -0.5f);
	v = (v >= -1) ? v : -1;
	v = (v <= +1) ? v : +1;
	return int(v * scale + round);
}

} // namespace
} // namespace vkvh

extern "C" {

size_t vkvh_meshlets_bound(size_t index_count, size_t max_vertices, size_t max_triangles) { // clusterizer.cpp:513-533
	const size_t conservative = max_vertices - 2;
	const size_t byVertices = (index_count + conservative - 1) / conservative;
	const size_t byTriangles = (index_count / 3 + max_triangles - 1) / max_triangles;
	return byVertices > byTriangles ? byVertices : byTriangles;
}

size_t vkvh_meshlets_build(vkvh_meshopt_Meshlet* meshlets, unsigned* meshlet_vertices, unsigned char* meshlet_triangles, const unsigned* indices,
                           size_t index_count, const float* vertex_positions, size_t vertex_count, size_t vertex_positions_stride,
                           size_t max_vertices, size_t max_triangles, float cone_weight) {
	if (index_count % 3 || vertex_positions_stride < 12 || vertex_positions_stride % 4 || max_vertices < 3 || max_vertices > 255 || max_triangles < 1 ||
	    max_triangles > 512 || max_triangles % 4 || !(cone_weight >= 0 && cone_weight <= 1))
		return 0;
	for (size_t i = 0; i < index_count; ++i)
		if (indices[i] >= vertex_count) return 0; // the library asserts; untrusted glTF gets a refusal instead
	vkvh::GreedyBuilder b(indices, index_count, vertex_positions, vertex_count, vertex_positions_stride, max_vertices, max_triangles, cone_weight);
	return b.run(meshlets, meshlet_vertices, meshlet_triangles);
}

// meshopt_optimizeMeshlet (clusterizer.cpp:886-974): order the triangles strip-like (prefer a triangle sharing >= 2 vertices with
// the last three emitted), then renumber the vertices in first-use order.
void vkvh_meshlet_optimize(unsigned* meshlet_vertices, unsigned char* meshlet_triangles, size_t triangle_count, size_t vertex_count) {
	if (triangle_count > 512 || vertex_count > 255) return;
	unsigned char* tri = meshlet_triangles;
	unsigned char stamp[255];
	std::memset(stamp, 0, vertex_count);
	unsigned char now = 128; // "nothing recent": every distance (now - stamp) starts at 128 >= window
	const unsigned char window = 3;
	for (size_t i = 0; i < triangle_count; ++i) {
		int pick = -1, pickHits = -1;
		for (size_t j = i; j < triangle_count; ++j) {
			const int hits = ((unsigned char)(now - stamp[tri[j * 3]]) < window) + ((unsigned char)(now - stamp[tri[j * 3 + 1]]) < window) +
			                 ((unsigned char)(now - stamp[tri[j * 3 + 2]]) < window);
			if (hits > pickHits) {
				pick = (int)j; pickHits = hits;
				if (pickHits >= 2) break;
			}
		}
		const unsigned char a = tri[pick * 3], b = tri[pick * 3 + 1], c = tri[pick * 3 + 2];
		std::memmove(tri + (i + 1) * 3, tri + i * 3, (size_t)(pick - (int)i) * 3); // keep the skipped triangles in order
		tri[i * 3] = a; tri[i * 3 + 1] = b; tri[i * 3 + 2] = c;
		++now;
		stamp[a] = now; stamp[b] = now; stamp[c] = now;
	}
	unsigned order[255];
	unsigned char remap[255];
	std::memset(remap, 0xff, vertex_count);
	size_t used = 0;
	for (size_t i = 0; i < triangle_count * 3; ++i) {
		unsigned char& r = remap[tri[i]];
		if (r == 0xff) { r = (unsigned char)used; order[used++] = meshlet_vertices[tri[i]]; }
		tri[i] = r;
	}
	std::memcpy(meshlet_vertices, order, used * sizeof(unsigned));
}

// meshopt_computeMeshletBounds (clusterizer.cpp:712-884): bounding sphere + normal cone of one meshlet.
void vkvh_meshlet_bounds(const unsigned* meshlet_vertices, const unsigned char* meshlet_triangles, size_t triangle_count, const float* vertex_positions,
                         size_t vertex_count, size_t vertex_positions_stride, vkvh_meshopt_Bounds* out) {
	std::memset(out, 0, sizeof(*out));
	if (triangle_count > 512) return;
	(void)vertex_count;
	const size_t strideF = vertex_positions_stride / sizeof(float);
	static thread_local float normals[512][3];
	static thread_local float corners[512][3][3];
	size_t n = 0;
	for (size_t t = 0; t < triangle_count; ++t) {
		const float* p0 = vertex_positions + strideF * meshlet_vertices[meshlet_triangles[t * 3]];
		const float* p1 = vertex_positions + strideF * meshlet_vertices[meshlet_triangles[t * 3 + 1]];
		const float* p2 = vertex_positions + strideF * meshlet_vertices[meshlet_triangles[t * 3 + 2]];
		const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
		const float e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
		const float nx = e1[1] * e2[2] - e1[2] * e2[1];
		const float ny = e1[2] * e2[0] - e1[0] * e2[2];
		const float nz = e1[0] * e2[1] - e1[1] * e2[0];
		const float area = sqrtf(nx * nx + ny * ny + nz * nz);
		if (area == 0.f) continue; // degenerate triangles are invisible anyway
		normals[n][0] = nx / area; normals[n][1] = ny / area; normals[n][2] = nz / area;
		std::memcpy(corners[n][0], p0, 12); std::memcpy(corners[n][1], p1, 12); std::memcpy(corners[n][2], p2, 12);
		++n;
	}
	if (n == 0) return; // no valid triangle: all-zero bounds = trivially rejected
	float ps[4] = {0, 0, 0, 0}, ns[4] = {0, 0, 0, 0};
	vkvh::boundingSphere(ps, corners[0], n * 3);
	vkvh::boundingSphere(ns, normals, n); // the normals as points: the sphere's centre is the cone axis
	float axis[3] = {ns[0], ns[1], ns[2]};
	const float len = sqrtf(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
	const float inv = len == 0.f ? 0.f : 1.f / len;
	for (float& a : axis) a *= inv;
	float mindp = 1.f;
	for (size_t i = 0; i < n; ++i) {
		const float dp = normals[i][0] * axis[0] + normals[i][1] * axis[1] + normals[i][2] * axis[2];
		mindp = (dp < mindp) ? dp : mindp;
	}
	out->center[0] = ps[0]; out->center[1] = ps[1]; out->center[2] = ps[2];
	out->radius = ps[3];
	if (mindp <= 0.1f) { // normals spread over ~168 degrees or more: the cone cannot reject anything
		out->cone_cutoff = 1;
		out->cone_cutoff_s8 = 127;
		return;
	}
	float maxt = 0;
	for (size_t i = 0; i < n; ++i) { // apex: the point on centre - t*axis behind every triangle's plane
		const float cx = ps[0] - corners[i][0][0], cy = ps[1] - corners[i][0][1], cz = ps[2] - corners[i][0][2];
		const float dc = cx * normals[i][0] + cy * normals[i][1] + cz * normals[i][2];
		const float dn = axis[0] * normals[i][0] + axis[1] * normals[i][1] + axis[2] * normals[i][2];
		const float t = dc / dn;
		maxt = (t > maxt) ? t : maxt;
	}
	for (int k = 0; k < 3; ++k) { out->cone_apex[k] = ps[k] - axis[k] * maxt; out->cone_axis[k] = axis[k]; }
	out->cone_cutoff = sqrtf(1 - mindp * mindp);
	for (int k = 0; k < 3; ++k) out->cone_axis_s8[k] = (signed char)vkvh::quantizeSnorm8(out->cone_axis[k]);
	const float e0 = fabsf(out->cone_axis_s8[0] / 127.f - out->cone_axis[0]);
	const float e1 = fabsf(out->cone_axis_s8[1] / 127.f - out->cone_axis[1]);
	const float e2 = fabsf(out->cone_axis_s8[2] / 127.f - out->cone_axis[2]);
	// (meshoptimizer converts this float unconditionally, clusterizer.cpp:858; positions with NaN / inf — a damaged asset — make it NaN, whose
	// conversion is undefined: those meshlets get the "never reject" cutoff)
	const float cutf = 127 * (out->cone_cutoff + e0 + e1 + e2) + 1;
	const int cut = (cutf < 128.f) ? int(cutf) : 128;
	out->cone_cutoff_s8 = (cut > 127) ? 127 : (signed char)cut;
}

} // extern "C"
