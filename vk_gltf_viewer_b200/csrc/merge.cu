// merge.cu — multi-GPU merge of the 64-bit visibility buffers: element-wise unsigned min over the ranks'
// W*H keys, in place, so that every rank ends up with the image a single GPU would have produced from the whole draw list.
// The reference is single-GPU; this is the one exchange step of the meshlet-range-sharded path (BASELINE config 5).
//
// One process per GPU; every rank maps its peers' visbuffers through CUDA IPC (ctx.cu: vkv_ipc_*), so the kernels below
// read and write peer HBM directly over NVLink / NVSwitch:
//   barrier  (1 block)   every rank's raster pass is complete and visible
//   merge    (grid)      fused reduce-scatter + all-gather: rank r owns strip r of the image, loads that strip from all
//                        ranks with 16-byte loads, takes the min and stores the result into strip r of every rank whose
//                        value differs (P2P stores) — no staging buffer, each key crosses NVLink once in and at most
//                        once out per peer; untouched (all-clear) regions and each pixel's winner cost no store
//   barrier  (1 block)   all strips have landed everywhere
// A key is (~depthBits << 32 | id): min == nearest fragment, ties == lowest id, exactly what atomicMin does inside one GPU,
// so the merged image is bit-identical to the single-GPU one.
#include "kernels.cuh"
#include "xgpu.cuh"

namespace {

// Thread t signals rank t (writes `epoch` into slot [rank] of rank t's flag array) and waits for rank t's signal in the
// local array (xgpu.cuh).
__global__ void xgpu_barrier_kernel(const MergeParams p, uint32_t epoch, unsigned long long timeout_ns) {
	const int t = threadIdx.x;
	if (t >= p.nranks) return;
	xgpu_signal(p.flags, p.rank, t, epoch);
	xgpu_wait(p.flags[p.rank], t, epoch, timeout_ns, p.error);
}

__device__ __forceinline__ ulonglong2 min2(ulonglong2 a, ulonglong2 b) {
	return make_ulonglong2(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y);
}

__device__ __forceinline__ bool differs(ulonglong2 a, ulonglong2 b) { return a.x != b.x || a.y != b.y; }

template <int N>
__global__ void __launch_bounds__(256) merge_min_kernel(const MergeParams p) {
	// strip r = pairs [n2*r/N, n2*(r+1)/N) of the image viewed as 16-byte pairs (the odd last key, if any, belongs to the last rank)
	constexpr int U = N <= 2 ? 4 : (N <= 4 ? 2 : 1); // pairs per thread per iteration: N*U 16-byte loads in flight
	const size_t n2 = p.n >> 1;
	const size_t lo = n2 * (size_t)p.rank / N, hi = n2 * (size_t)(p.rank + 1) / N;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i0 = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * U) {
		ulonglong2 v[U][N];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const size_t i = i0 + u * stride;
#pragma unroll
			for (int k = 0; k < N; ++k) { // peer order rotated by rank: at any moment the ranks pull from different peers
				const int r = (p.rank + k) % N;
				if (i < hi) v[u][k] = __ldcg((const ulonglong2*)p.vis[r] + i);
			}
		}
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const size_t i = i0 + u * stride;
			if (i >= hi) continue;
			ulonglong2 m = v[u][0];
#pragma unroll
			for (int k = 1; k < N; ++k) m = min2(m, v[u][k]);
#pragma unroll
			for (int k = 0; k < N; ++k) { // a rank that already holds the minimum (or where every rank is still clear) gets no store
				const int r = (p.rank + k) % N;
				if (differs(m, v[u][k])) __stcg((ulonglong2*)p.vis[r] + i, m);
			}
		}
	}
	if ((p.n & 1) && p.rank == N - 1 && blockIdx.x == 0 && threadIdx.x == 0) {
		unsigned long long m = p.vis[0][p.n - 1];
		for (int r = 1; r < N; ++r) { const unsigned long long x = p.vis[r][p.n - 1]; m = x < m ? x : m; }
		for (int r = 0; r < N; ++r) p.vis[r][p.n - 1] = m;
	}
}

// generic rank count (not 2/4/8): same algorithm, runtime loop
__global__ void __launch_bounds__(256) merge_min_generic_kernel(const MergeParams p) {
	const int N = p.nranks;
	const size_t n2 = p.n >> 1;
	const size_t lo = n2 * (size_t)p.rank / N, hi = n2 * (size_t)(p.rank + 1) / N;
	for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
		ulonglong2 m = __ldcg((const ulonglong2*)p.vis[0] + i);
		for (int r = 1; r < N; ++r) m = min2(m, __ldcg((const ulonglong2*)p.vis[r] + i));
		for (int r = 0; r < N; ++r) __stcg((ulonglong2*)p.vis[r] + i, m);
	}
	if ((p.n & 1) && p.rank == N - 1 && blockIdx.x == 0 && threadIdx.x == 0) {
		unsigned long long m = p.vis[0][p.n - 1];
		for (int r = 1; r < N; ++r) { const unsigned long long x = p.vis[r][p.n - 1]; m = x < m ? x : m; }
		for (int r = 0; r < N; ++r) p.vis[r][p.n - 1] = m;
	}
}

} // namespace

cudaError_t launch_xgpu_barrier(const MergeParams& p, uint32_t epoch, unsigned long long timeout_ns, cudaStream_t stream) {
	xgpu_barrier_kernel<<<1, 32, 0, stream>>>(p, epoch, timeout_ns);
	return cudaGetLastError();
}

cudaError_t launch_merge_min(const MergeParams& p, int num_sms, cudaStream_t stream) {
	const size_t strip = (p.n / 2) / (size_t)p.nranks + 1;
	size_t grid = (strip + 255) / 256;
	if (grid > (size_t)num_sms * 8) grid = (size_t)num_sms * 8;
	if (grid == 0) grid = 1;
	switch (p.nranks) {
		case 2: merge_min_kernel<2><<<(unsigned)grid, 256, 0, stream>>>(p); break;
		case 4: merge_min_kernel<4><<<(unsigned)grid, 256, 0, stream>>>(p); break;
		case 8: merge_min_kernel<8><<<(unsigned)grid, 256, 0, stream>>>(p); break;
		default: merge_min_generic_kernel<<<(unsigned)grid, 256, 0, stream>>>(p); break;
	}
	return cudaGetLastError();
}
