// meshlets.cu — meshlet partition + bounds on the device.
// Stands where PrimitiveProcessingTask::processPrimitive builds meshlets and their AABBs on the host
// (src/vk_gltf_viewer/assets.cpp:322-373: meshopt_buildMeshlets, then per meshlet meshopt_optimizeMeshlet + min/max bounds).
// The partition produced here is meshoptimizer's SCAN partition (submodules/meshoptimizer/src/clusterizer.cpp:224-298,673-708,
// meshopt_buildMeshletsScan): triangles in index-buffer order, a meshlet closes when the next triangle's unused corners would
// exceed max_vertices or it already holds max_triangles — byte-identical to that function (tests/test_gpu_meshlets.py against
// vectors produced by the reference's own library).  The reference calls the kd-tree / adjacency variant, which has no golden
// output and no parallel form; any valid partition renders the same image.
//
// The scan rule looks sequential (where meshlet k+1 starts depends on where meshlet k ended), but only its CHAIN is:
//   len      one thread per triangle s answers "if a meshlet started at s, how many triangles would it take, how many distinct
//            vertices?" — a window walk with a 128-slot hash set in local memory; all starts at once, ~1 % of the answers are
//            used, and it is still the cheap way to buy parallelism (20 M triangles: a few ms);
//   segment  the triangle range is cut into segments of kMeshletSeg triangles; because a meshlet is at most max_triangles long, the
//            chain enters a segment within its first max_triangles triangles: one thread per (segment, entry) follows
//            s -> s + len[s] through the segment and notes where it leaves and what it produced (meshlets, vertices, bytes);
//   chain    ONE thread strings the segments together (a table lookup per segment: 10 k steps for 20 M triangles) and leaves
//            every segment's entry point and output bases; primitives restart the chain;
//   walk     one thread per segment replays its part of the chain and writes one record per meshlet (start, output offsets);
//   emit     one thread per meshlet: local vertex numbering in first-appearance order, triangle bytes (+ zero padding to 4),
//            bounds exactly as assets.cpp:349-372 (glm::min / glm::max from the first vertex, center = (min+max)*0.5,
//            extents = max - center), and the reference's 36-byte Meshlet record.
#include "kernels.cuh"

namespace {

constexpr uint32_t kSetSlots = 128;    // hash-set slots per thread (max_vertices <= 64 (+2 transient) entries)
constexpr uint32_t kEmpty = 0xffffffffu;

__device__ __forceinline__ uint32_t slot_of(uint32_t v) { return (v * 2654435761u) >> 25; } // 7 bits

// last index with first[i] <= x (first[] ascending, first[0] == 0)
__device__ __forceinline__ uint32_t find_owner(const uint32_t* __restrict__ first, uint32_t n, uint32_t x) {
	uint32_t lo = 0, hi = n - 1;
	while (lo < hi) {
		const uint32_t mid = (lo + hi + 1) >> 1;
		if (first[mid] <= x) lo = mid; else hi = mid - 1;
	}
	return lo;
}

struct VertexSet { // open addressing, linear probing
	uint32_t key[kSetSlots];
	__device__ void clear() {
#pragma unroll 8
		for (uint32_t i = 0; i < kSetSlots; ++i) key[i] = kEmpty;
	}
	__device__ bool has(uint32_t v) const {
		for (uint32_t s = slot_of(v);; s = (s + 1) & (kSetSlots - 1)) {
			const uint32_t k = key[s];
			if (k == v) return true;
			if (k == kEmpty) return false;
		}
	}
	__device__ uint32_t put(uint32_t v) { // v must be absent; returns its slot
		uint32_t s = slot_of(v);
		while (key[s] != kEmpty) s = (s + 1) & (kSetSlots - 1);
		key[s] = v;
		return s;
	}
	__device__ int slot(uint32_t v) const { // -1 when absent
		for (uint32_t s = slot_of(v);; s = (s + 1) & (kSetSlots - 1)) {
			const uint32_t k = key[s];
			if (k == v) return (int)s;
			if (k == kEmpty) return -1;
		}
	}
};

// ---- len: the greedy meshlet that would start at every triangle -------------------------------------------------------------
__global__ void __launch_bounds__(128) mb_len_kernel(const MeshletBuildPrim* __restrict__ prims, uint32_t nPrims, const uint32_t* __restrict__ triFirst,
                                                     uint32_t totalTris, uint32_t maxV, uint32_t maxT, uint8_t* __restrict__ len, uint8_t* __restrict__ ucnt,
                                                     uint32_t* __restrict__ badIndex) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= totalTris) return;
	const uint32_t p = find_owner(triFirst, nPrims, s);
	const uint32_t* __restrict__ idx = prims[p].indices;
	const uint32_t nT = triFirst[p + 1] - triFirst[p];
	{ // every thread range-checks ITS triangle: the emit pass dereferences vertices[index] (untrusted glTF; clusterizer.cpp:45 asserts)
		const uint32_t nV = prims[p].vertexCount, t = s - triFirst[p];
		if (nV && (__ldg(idx + 3 * t) >= nV || __ldg(idx + 3 * t + 1) >= nV || __ldg(idx + 3 * t + 2) >= nV)) atomicMax(badIndex, p + 1);
	}
	VertexSet set;
	set.clear();
	uint32_t vc = 0, tc = 0;
	for (uint32_t t = s - triFirst[p]; t < nT; ++t) {
		const uint32_t a = __ldg(idx + 3 * t), b = __ldg(idx + 3 * t + 1), c = __ldg(idx + 3 * t + 2);
		const bool ha = set.has(a), hb = set.has(b), hc = set.has(c);
		// appendMeshlet (clusterizer.cpp:233-246): unused corners are counted per corner, a repeated corner counts twice
		const uint32_t extra = (ha ? 0u : 1u) + (hb ? 0u : 1u) + (hc ? 0u : 1u);
		if (vc + extra > maxV || tc >= maxT) break;
		if (!ha) { set.put(a); ++vc; }
		if (!hb && b != a) { set.put(b); ++vc; }
		if (!hc && c != a && c != b) { set.put(c); ++vc; }
		++tc;
	}
	len[s] = (uint8_t)tc;
	ucnt[s] = (uint8_t)vc;
}

__device__ __forceinline__ uint32_t padded_bytes(uint32_t tris) { return (tris * 3 + 3) & ~3u; }

// ---- segment: for every way the chain can enter a segment, where it leaves and what it produces --------------------------
__global__ void mb_segment_kernel(const MeshletBuildSeg* __restrict__ segs, uint32_t nSegs, uint32_t maxT, const uint8_t* __restrict__ len,
                                  const uint8_t* __restrict__ ucnt, MeshletBuildSegEntry* __restrict__ table) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nSegs * maxT) return;
	const uint32_t g = i / maxT, x = i % maxT;
	const MeshletBuildSeg sg = segs[g];
	MeshletBuildSegEntry e = {0, 0, 0, 0};
	uint32_t s = sg.start + x;
	while (s < sg.end) {
		const uint32_t n = len[s];
		e.meshlets += 1; e.vertices += ucnt[s]; e.bytes += padded_bytes(n);
		s += n;
	}
	e.exit = s >= sg.end ? s - sg.end : 0; // offset into the next segment (meaningless past a primitive's last segment)
	table[i] = e;
}

// ---- chain: string the segments together (serial by nature; one lookup per segment) ------------------------------------------
__global__ void mb_chain_kernel(const MeshletBuildSeg* __restrict__ segs, uint32_t nSegs, uint32_t maxT, const MeshletBuildSegEntry* __restrict__ table,
                                MeshletBuildSegState* __restrict__ state, uint32_t* __restrict__ primBase /* [nPrims + 1][3] */, uint32_t nPrims) {
	if (blockIdx.x || threadIdx.x) return;
	uint32_t M = 0, V = 0, B = 0, entry = 0, prim = 0;
	for (uint32_t g = 0; g < nSegs; ++g) {
		const MeshletBuildSeg sg = segs[g];
		if (sg.first_of_prim) {
			entry = 0;
			for (; prim <= sg.prim; ++prim) { primBase[prim * 3] = M; primBase[prim * 3 + 1] = V; primBase[prim * 3 + 2] = B; } // empty primitives too
		}
		state[g] = MeshletBuildSegState{entry, M, V, B};
		const MeshletBuildSegEntry e = table[g * maxT + entry];
		M += e.meshlets; V += e.vertices; B += e.bytes;
		entry = e.exit;
	}
	for (; prim <= nPrims; ++prim) { primBase[prim * 3] = M; primBase[prim * 3 + 1] = V; primBase[prim * 3 + 2] = B; }
}

// ---- walk: one record per meshlet -----------------------------------------------------------------------------------------------
__global__ void mb_walk_kernel(const MeshletBuildSeg* __restrict__ segs, uint32_t nSegs, const MeshletBuildSegState* __restrict__ state,
                               const uint8_t* __restrict__ len, const uint8_t* __restrict__ ucnt, MeshletBuildRecord* __restrict__ rec) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nSegs) return;
	const MeshletBuildSeg sg = segs[g];
	const MeshletBuildSegState st = state[g];
	uint32_t m = st.meshlets, v = st.vertices, b = st.bytes;
	for (uint32_t s = sg.start + st.entry; s < sg.end;) {
		const uint32_t n = len[s];
		rec[m] = MeshletBuildRecord{s, v, b};
		++m; v += ucnt[s]; b += padded_bytes(n);
		s += n;
	}
}

// ---- emit: vertex numbering, triangle bytes, bounds, Meshlet record ------------------------------------------------------------
__global__ void __launch_bounds__(128) mb_emit_kernel(const MeshletBuildPrim* __restrict__ prims, uint32_t nPrims, const uint32_t* __restrict__ triFirst,
                                                      const uint32_t* __restrict__ primBase, const MeshletBuildRecord* __restrict__ rec, uint32_t nMeshlets,
                                                      const uint8_t* __restrict__ len, uint32_t vertexStride, vkv_Meshlet* __restrict__ meshlets,
                                                      uint32_t* __restrict__ meshletVertices, uint8_t* __restrict__ meshletTriangles) {
	const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= nMeshlets) return;
	const MeshletBuildRecord r = rec[m];
	const uint32_t p = find_owner(triFirst, nPrims, r.start);
	const MeshletBuildPrim pr = prims[p];
	const uint32_t n = len[r.start];
	VertexSet set;
	uint8_t local[kSetSlots]; // local vertex number per hash slot
	set.clear();
	uint32_t vc = 0;
	float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
	uint32_t* mv = meshletVertices + r.vertices;
	uint8_t* mt = meshletTriangles + r.bytes;
	const uint32_t t0 = r.start - triFirst[p];
	for (uint32_t t = 0; t < n; ++t) {
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const uint32_t v = __ldg(pr.indices + 3 * (t0 + t) + k);
			int s = set.slot(v);
			if (s < 0) {
				s = (int)set.put(v);
				local[s] = (uint8_t)vc;
				mv[vc] = v;
				const float* q = (const float*)(pr.vertices + (size_t)v * vertexStride);
				const float x[3] = {__ldg(q), __ldg(q + 1), __ldg(q + 2)};
#pragma unroll
				for (int c = 0; c < 3; ++c) {
					if (vc == 0) { mn[c] = mx[c] = x[c]; }
					else { mn[c] = gmin(mn[c], x[c]); mx[c] = gmax(mx[c], x[c]); } // glm::min / glm::max (assets.cpp:361-362)
				}
				++vc;
			}
			mt[t * 3 + k] = local[s];
		}
	}
	for (uint32_t o = n * 3; o & 3u; ++o) mt[o] = 0; // finishMeshlet: zero padding to 4 bytes
	// the reference's 36-byte Meshlet (mesh_common.h.glsl:49-58): vertexOffset, triangleOffset, vertexCount u8, triangleCount u8,
	// two padding bytes (written as zero), aabbExtents, aabbCenter
	uint32_t w[9];
	w[0] = r.vertices - primBase[p * 3 + 1];
	w[1] = r.bytes - primBase[p * 3 + 2];
	w[2] = vc | (n << 8);
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		const float ctr = (mn[c] + mx[c]) * 0.5f;
		w[3 + c] = __float_as_uint(mx[c] - ctr);
		w[6 + c] = __float_as_uint(ctr);
	}
	uint32_t* dst = (uint32_t*)(meshlets + m);
#pragma unroll
	for (int i = 0; i < 9; ++i) dst[i] = w[i];
}

} // namespace

cudaError_t launch_meshlet_scan(const MeshletBuildJob& j, cudaStream_t stream) {
	if (j.totalTris == 0) return cudaSuccess;
	mb_len_kernel<<<(j.totalTris + 127) / 128, 128, 0, stream>>>(j.prims, j.nPrims, j.triFirst, j.totalTris, j.maxV, j.maxT, j.len, j.ucnt, j.badIndex);
	const uint32_t entries = j.nSegs * j.maxT;
	mb_segment_kernel<<<(entries + 127) / 128, 128, 0, stream>>>(j.segs, j.nSegs, j.maxT, j.len, j.ucnt, j.table);
	mb_chain_kernel<<<1, 32, 0, stream>>>(j.segs, j.nSegs, j.maxT, j.table, j.state, j.primBase, j.nPrims);
	return cudaGetLastError();
}

cudaError_t launch_meshlet_emit(const MeshletBuildJob& j, uint32_t nMeshlets, vkv_Meshlet* meshlets, uint32_t* meshletVertices, uint8_t* meshletTriangles,
                                cudaStream_t stream) {
	if (nMeshlets == 0) return cudaSuccess;
	mb_walk_kernel<<<(j.nSegs + 127) / 128, 128, 0, stream>>>(j.segs, j.nSegs, j.state, j.len, j.ucnt, j.rec);
	mb_emit_kernel<<<(nMeshlets + 127) / 128, 128, 0, stream>>>(j.prims, j.nPrims, j.triFirst, j.primBase, j.rec, nMeshlets, j.len, j.vertexStride, meshlets,
	                                                           meshletVertices, meshletTriangles);
	return cudaGetLastError();
}
