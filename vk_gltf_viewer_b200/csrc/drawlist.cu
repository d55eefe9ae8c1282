// drawlist.cu — device-side MeshletDraw[] generation.
// Replaces World::rebuildDrawBuffer (world.cpp:230-293): the reference walks the node hierarchy on the host, emits one
// MeshletDraw{primitive, i, transformIndex} per (mesh-node, primitive, meshlet) into a std::vector and memcpy's it into a
// mapped buffer ("often multiple milliseconds", world.cpp:284-285; 12 B per draw = 12.7 MB at config 3, 129 MB at config 5).
// Here the host uploads only the SEGMENTS — one {primitiveIndex, transformIndex} per (mesh-node, primitive) in traversal
// order, 8 B each — and the list is expanded on the device:
//   1. segment lengths = Primitive.meshletCount, exclusive scan (one block, 1024-wide chunks with a running carry);
//   2. one warp per segment writes its draws with coalesced stores.
// The result is byte-identical to the host-built list (tests/test_gpu_parity.py::test_device_draw_list).
#include "kernels.cuh"

namespace {

__global__ void __launch_bounds__(1024) segment_scan_kernel(const vkv_DrawSegment* __restrict__ seg, uint32_t n, const vkv_Primitive* __restrict__ prims,
                                                            uint32_t* __restrict__ offsets /* n + 1 */, uint32_t* __restrict__ overflow) {
	__shared__ uint32_t warpSums[32];
	__shared__ unsigned long long carry;
	if (threadIdx.x == 0) carry = 0ull;
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (uint32_t base = 0; base < n; base += 1024) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t len = i < n ? __ldg(&prims[__ldg(&seg[i].primitiveIndex)].meshletCount) : 0u;
		uint32_t v = len; // inclusive warp scan
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
			if (lane >= (uint32_t)d) v += t;
		}
		if (lane == 31) warpSums[warp] = v;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = warpSums[lane];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t t = __shfl_up_sync(0xffffffffu, w, d);
				if (lane >= (uint32_t)d) w += t;
			}
			warpSums[lane] = w; // inclusive over warps
		}
		__syncthreads();
		const unsigned long long c = carry;
		const unsigned long long excl = c + (warp ? warpSums[warp - 1] : 0u) + (v - len);
		if (i < n) offsets[i] = (uint32_t)excl;
		__syncthreads();
		if (threadIdx.x == 1023) {
			const unsigned long long total = c + warpSums[31];
			carry = total;
			if (total > VKV_MAX_MESHLET_DRAWS) *overflow = 1u; // visbuffer.h.glsl:15-17: 25-bit drawIndex
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) offsets[n] = (uint32_t)(carry > 0xffffffffull ? 0xffffffffull : carry);
}

__global__ void __launch_bounds__(256) expand_segments_kernel(const vkv_DrawSegment* __restrict__ seg, uint32_t n, const uint32_t* __restrict__ offsets,
                                                              vkv_MeshletDraw* __restrict__ draws, uint32_t capacity) {
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warpsPerGrid = gridDim.x * (blockDim.x >> 5);
	for (uint32_t s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n; s += warpsPerGrid) {
		const uint32_t lo = __ldg(offsets + s), hi = __ldg(offsets + s + 1);
		const uint32_t prim = __ldg(&seg[s].primitiveIndex), xf = __ldg(&seg[s].transformIndex);
		// world.cpp:256-262: for i in [0, meshletCount): MeshletDraw{primitive, i, transformIndex}; written as a flat u32 stream so
		// consecutive lanes store consecutive words
		uint32_t* out = (uint32_t*)(draws + lo);
		const uint32_t words = (min(hi, capacity) > lo ? min(hi, capacity) - lo : 0u) * 3u;
		for (uint32_t w = lane; w < words; w += 32) {
			const uint32_t i = w / 3u, f = w - i * 3u;
			out[w] = f == 0 ? prim : (f == 1 ? i : xf);
		}
	}
}

} // namespace

cudaError_t launch_segment_scan(const vkv_DrawSegment* seg, uint32_t n, const vkv_Primitive* prims, uint32_t* offsets, uint32_t* overflow, cudaStream_t stream) {
	segment_scan_kernel<<<1, 1024, 0, stream>>>(seg, n, prims, offsets, overflow);
	return cudaGetLastError();
}

cudaError_t launch_expand_segments(const vkv_DrawSegment* seg, uint32_t n, const uint32_t* offsets, vkv_MeshletDraw* draws, uint32_t capacity,
                                   int num_sms, cudaStream_t stream) {
	if (!n) return cudaSuccess;
	uint32_t grid = (n + 7) / 8;
	if (grid > (uint32_t)num_sms * 8) grid = (uint32_t)num_sms * 8;
	expand_segments_kernel<<<grid, 256, 0, stream>>>(seg, n, offsets, draws, capacity);
	return cudaGetLastError();
}
