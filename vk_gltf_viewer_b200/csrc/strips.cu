// strips.cu — the exchange step of the meshlet-range-sharded path, by screen-strip ownership ("cheaper variant").
//
// Every rank culls and rasterises ITS range of the draw list into its own full-resolution 64-bit visbuffer (raster.cu marks
// every 64x16-pixel tile a drawn triangle's bounding box touches).  The screen is cut into rows of tiles (16 pixel rows each) dealt
// round-robin: rank r OWNS tile rows r, r + n, r + 2n, … ("its strips").  ONE kernel per rank does, for every tile it owns:
//   reduce-scatter   read the peers' dirty flags of the tile, pull only the tiles some peer actually drew into (16-byte loads
//                    over NVLink / NVSwitch peer memory), min them into the owner's keys, store the merged rows back locally;
//   HiZ per strip    the exact-2x mips of the merged tile, straight from the registers that hold it (hiz_tile.cuh);
//   all-gather       every mip texel that CHANGED is stored into every rank's pyramid (peer stores).
// After a second barrier each rank rebuilds the few small mips locally from the all-gathered last exact mip (hiz.cu's tail).
// The full-resolution keys never travel whole: per frame a rank receives the tiles its peers touched inside its strip
// (against 2*(n-1)/n * 8*W*H bytes for an all-reduce of the visbuffer) and sends 4 bytes per changed pyramid texel per peer.
// The merged visbuffer stays distributed — rank r holds rows of strip r — which is what a per-strip consumer (resolve) wants;
// vkv_gather_strips assembles the whole image on every rank for readers that need it.
//
// A key is (~depthBits << 32 | id): min == nearest fragment, ties == lowest id, exactly what atomicMin does inside one GPU, so
// strip r on rank r is bit-identical to the same rows of a single-GPU frame, and so is the pyramid on every rank.
// The reference is single-GPU: there is no reference call site for this file.
#include "hiz_tile.cuh"
#include "kernels.cuh"
#include "xgpu.cuh"

namespace {

constexpr int kStripThreads = 256;

__device__ __forceinline__ ulonglong2 min2(ulonglong2 a, ulonglong2 b) { return make_ulonglong2(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y); }

// Persistent blocks (four per SM), each owning a contiguous range of the tiles of this rank's strip, one warp per tile.
//   0. the block fetches, for ALL its tiles at once, which ranks drew into them (one NVLink round trip per block);
//   1. merge: one warp per (tile, peer that drew into it) — the peer's rows (16-byte loads, 8 in flight per lane) are reduced into
//      the owner's keys with 64-bit atomic mins, all pairs of the block in flight together;
//   2. pyramid: the merged tile's 16 rows AND the 15 pyramid texels this lane may produce are requested together (one round trip),
//      the tile is reduced in registers (hiz_tile.cuh), and every texel that differs from the local pyramid is stored into
//      every rank's pyramid.
// Keeping 1 and 2 apart (instead of holding the 16 merged rows in registers across both) keeps the kernel at 64 registers: the
// strip is read from DRAM (an 8K visbuffer is twice the L2), and it takes 32 warps per SM with a tile each in flight to keep DRAM busy.
constexpr uint32_t kRound = 512; // tiles whose masks fit the shared-memory table at once

__global__ void __launch_bounds__(kStripThreads, 4) strip_merge_hiz_kernel(const StripParams p) {
	__shared__ uint32_t sMask[kRound]; // bit r: rank r drew into the tile in this pass (bit `me`: this rank did)
	__shared__ unsigned short sPair[kRound * (kMaxRanks - 1)]; // (tile of the round << 4) | peer, for every pair to pull
	__shared__ uint32_t sPairs;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int N = p.mp.nranks, me = p.mp.rank;
	const uint32_t nTiles = strip_tile_rows_owned(p.tilesY, me, N) * p.tilesX; // tile t of this rank: tile row me + N * (t / tilesX)
	// barrier in: this kernel starts when THIS rank's raster pass is complete (stream order); block 0 says so to every rank, and
	// every block waits until all ranks have said so (polling this rank's own slots)
	if (blockIdx.x == 0 && (int)threadIdx.x < N) xgpu_signal(p.mp.flags, me, threadIdx.x, p.epoch_in);
	if ((int)threadIdx.x < N) xgpu_wait(p.mp.flags[me], threadIdx.x, p.epoch_in, p.timeout_ns, p.mp.error);
	__syncthreads();
	const HizTileGeo geo = {p.W, p.H, p.exact_levels, {p.pyr.off[0], p.pyr.off[1], p.pyr.off[2], p.pyr.off[3]}, {p.pyr.w[0], p.pyr.w[1], p.pyr.w[2], p.pyr.w[3]}};
	unsigned long long* const vis = p.mp.vis[me];
	const float* const localPyr = p.mp.pyr[me];
	uint32_t pulled = 0, sent = 0; // this warp's tiles pulled over NVLink / this lane's texels stored to peers (statistics)
	const uint32_t perBlock = (nTiles + gridDim.x - 1) / gridDim.x;
	const uint32_t first = blockIdx.x * perBlock, last = min(nTiles, first + perBlock);
	for (uint32_t base = first; base < last; base += kRound) {
		const uint32_t cnt = min(kRound, last - base);
		__syncthreads(); // the previous round's readers of sMask are done
		for (uint32_t j = threadIdx.x; j < cnt; j += kStripThreads) sMask[j] = 0u;
		__syncthreads();
		// 0. dirty bytes of the whole round: work item (r, j) asks rank r about tile base + j (r == me: this rank's own marks, local)
		for (uint32_t w = threadIdx.x; w < cnt * (uint32_t)N; w += kStripThreads) {
			const uint32_t r = w / cnt, j = w % cnt, t = base + j;
			const uint32_t tile = ((uint32_t)me + (uint32_t)N * (t / p.tilesX)) * p.tilesX + t % p.tilesX;
			if (*(volatile const uint8_t*)(p.mp.dirty[r] + (size_t)p.pass * p.dirtyStride + tile)) atomicOr(&sMask[j], 1u << r);
		}
		__syncthreads();
		// 1. merge, one warp per (tile, peer that drew into it): the peer's rows (16-byte loads over NVLink, 8 in flight per lane) go into
		// the owner's keys with the same 64-bit atomic min the rasteriser uses — every pair of the block is in flight at once, where a
		// warp walking the peers of its tile one after the other would chain up to 2 x 7 NVLink round trips
		if (threadIdx.x == 0) sPairs = 0u;
		__syncthreads();
		for (uint32_t j = threadIdx.x; j < cnt; j += kStripThreads) {
			const uint32_t peers = sMask[j] & ~(1u << me);
			if (!peers) continue;
			uint32_t at = atomicAdd(&sPairs, (uint32_t)__popc(peers));
			for (uint32_t m = peers; m; m &= m - 1) sPair[at++] = (unsigned short)((j << 4) | (uint32_t)(__ffs(m) - 1));
		}
		__syncthreads();
		const uint32_t nPairs = sPairs;
		for (uint32_t u = warp; u < nPairs; u += kStripThreads / 32) {
			const uint32_t j = sPair[u] >> 4, r = sPair[u] & 15u, t = base + j;
			const uint32_t tx = t % p.tilesX, ty = (uint32_t)me + (uint32_t)N * (t / p.tilesX);
			const uint32_t x0 = tx * kTileW + lane * 2, y0 = ty * kTileH;
			if (x0 >= p.W) continue;
			const unsigned long long* pv = p.mp.vis[r];
#pragma unroll 1
			for (int half = 0; half < 2; ++half) {
				ulonglong2 q[8];
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					const uint32_t y = y0 + half * 8 + k;
					q[k] = make_ulonglong2(~0ull, ~0ull);
					if (y < p.H) q[k] = __ldcg((const ulonglong2*)(pv + (size_t)y * p.W + x0));
				}
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					unsigned long long* dst = vis + (size_t)(y0 + half * 8 + k) * p.W + x0;
					if (q[k].x != ~0ull) atomicMin(dst, q[k].x);      // clear keys cannot win
					if (q[k].y != ~0ull) atomicMin(dst + 1, q[k].y);
				}
			}
		}
		if (nPairs) __threadfence(); // this warp's reductions are performed before the block's tiles are re-read below
		__syncthreads();
		if (warp == 0 && lane == 0) pulled += nPairs;
		for (uint32_t j = warp; j < cnt; j += kStripThreads / 32) {
			const uint32_t t = base + j;
			const uint32_t tx = t % p.tilesX, ty = (uint32_t)me + (uint32_t)N * (t / p.tilesX);
			const uint32_t mask = sMask[j];
			// second pass of a frame: a tile neither a peer nor this rank drew into since the first exchange still holds the merged
			// keys the first exchange built its mips from — nothing to pull, nothing to rebuild
			if (p.pass == 1 && !mask) continue;
			const uint32_t x0 = tx * kTileW + lane * 2, y0 = ty * kTileH;
			const bool colIn = x0 < p.W;
			// 2. exact mips of the merged tile -> every rank's pyramid, changed texels only (all pyramids are identical before this
			// frame's stores, so the local copy tells whether a texel changes anywhere).  Which texels this lane owns depends on the
			// tile and the lane only, so their old values are requested together with the tile's rows.
			ulonglong2 v[kTileH];
#pragma unroll
			for (int r = 0; r < kTileH; ++r) {
				v[r] = make_ulonglong2(0ull, 0ull);
				if (colIn && y0 + r < p.H) v[r] = __ldcg((const ulonglong2*)(vis + (size_t)(y0 + r) * p.W + x0)); // L2: this warp may just have written it
			}
			uint32_t idxOf[kTileSlots], oldOf[kTileSlots];
			uint32_t have = 0; // bit s: this lane owns a texel in slot s
			hiz_tile_slots(geo, tx, ty, lane, [&](int slot, uint32_t idx) { idxOf[slot] = idx; have |= 1u << slot; });
#pragma unroll
			for (int sl = 0; sl < kTileSlots; ++sl) oldOf[sl] = (have >> sl) & 1u ? __float_as_uint(__ldcg(localPyr + idxOf[sl])) : 0u;
			hiz_tile_reduce(v, geo, tx, ty, lane, [&](int slot, uint32_t idx, float m) {
				if (oldOf[slot] != __float_as_uint(m)) {
					for (int r = 0; r < N; ++r) __stcg(p.mp.pyr[r] + idx, m);
					sent += (uint32_t)(N - 1);
				}
			});
		}
	}
	if (p.stats) {
		for (int o = 16; o; o >>= 1) sent += __shfl_xor_sync(0xffffffffu, sent, o);
		if (lane == 0) {
			if (pulled) atomicAdd(p.stats, pulled); // non-zero in warp 0 only
			if (sent) atomicAdd(p.stats + 1, sent);
		}
	}
	// barrier out: the block that finishes last tells every rank that this rank's strip is merged and its mips are stored everywhere
	// (the small-mip tail that follows on each rank waits for all these signals)
	__shared__ uint32_t sLast;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system(); // this block's peer stores before the ticket
		const uint32_t ticket = atomicAdd(p.done, 1u);
		sLast = (ticket == gridDim.x - 1) ? 1u : 0u;
		if (sLast) *p.done = 0u;
	}
	__syncthreads();
	if (sLast && (int)threadIdx.x < N) xgpu_signal(p.mp.flags, me, threadIdx.x, p.epoch_out);
}

// all-gather of the merged strips: every rank pulls the strips it does not own from their owners (readers that want the whole
// image on one GPU: parity tests, vkv_read_visbuffer64 after a strip-mode frame)
__global__ void __launch_bounds__(256) strip_gather_kernel(const StripParams p) {
	const int N = p.mp.nranks, me = p.mp.rank;
	ulonglong2* const vis = (ulonglong2*)p.mp.vis[me];
	const size_t rowPairs = p.W / 2, total = rowPairs * p.H; // W is even in strip mode (one exact mip at least)
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t y = (uint32_t)(i / rowPairs);
		const int owner = strip_owner_of_tile_row(y / kTileH, N);
		if (owner != me) vis[i] = __ldcg((const ulonglong2*)p.mp.vis[owner] + i);
	}
}

// order-independent 64-bit digest of data[first, first + count): sum over i of mix(data[i] + golden * (first + i)) mod 2^64
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}
__global__ void __launch_bounds__(256) hash64_kernel(const unsigned long long* __restrict__ data, size_t first, size_t count, unsigned long long* out) {
	unsigned long long acc = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
		acc += mix64(data[first + i] + 0x9e3779b97f4a7c15ull * (unsigned long long)(first + i + 1));
	for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// digest of the rows `rank` owns among `nranks` (position-dependent like hash64_kernel: equal rows at equal positions give equal sums)
__global__ void __launch_bounds__(256) hash_owned_kernel(const unsigned long long* __restrict__ vis, uint32_t W, uint32_t H, int rank, int nranks, unsigned long long* out) {
	unsigned long long acc = 0;
	const size_t total = (size_t)W * H;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t y = (uint32_t)(i / W);
		if (strip_owner_of_tile_row(y / kTileH, nranks) == rank) acc += mix64(vis[i] + 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1));
	}
	for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

} // namespace

cudaError_t launch_hash_owned(const unsigned long long* vis, uint32_t W, uint32_t H, int rank, int nranks, unsigned long long* out, int num_sms, cudaStream_t stream) {
	hash_owned_kernel<<<num_sms * 8, 256, 0, stream>>>(vis, W, H, rank, nranks, out);
	return cudaGetLastError();
}

cudaError_t launch_strip_merge_hiz(const StripParams& p, int num_sms, cudaStream_t stream) {
	const uint32_t tiles = strip_tile_rows_owned(p.tilesY, p.mp.rank, p.mp.nranks) * p.tilesX;
	if (tiles == 0) return cudaSuccess;
	uint32_t grid = (tiles + 7) / 8;                                 // at least a tile per warp
	if (grid > (uint32_t)num_sms * 4) grid = (uint32_t)num_sms * 4; // persistent: four blocks per SM (64 registers)
	strip_merge_hiz_kernel<<<grid, kStripThreads, 0, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_strip_gather(const StripParams& p, int num_sms, cudaStream_t stream) {
	strip_gather_kernel<<<num_sms * 4, 256, 0, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_hash64(const unsigned long long* data, size_t first, size_t count, unsigned long long* out, int num_sms, cudaStream_t stream) {
	if (count == 0) return cudaSuccess;
	size_t grid = (count + 255) / 256;
	if (grid > (size_t)num_sms * 8) grid = (size_t)num_sms * 8;
	hash64_kernel<<<(unsigned)grid, 256, 0, stream>>>(data, first, count, out);
	return cudaGetLastError();
}
