// ctx.cu — the extern "C" boundary (include/vkv.h): context, device memory, frame orchestration, readback.
// There is no CPU fallback anywhere in this file: without a CUDA device vkv_create fails with VKV_ERR_NO_DEVICE.
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <iterator>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>

#include <nvtx3/nvToolsExt.h>

#include "kernels.cuh"

struct vkv_ctx {
	int device = 0;
	int num_sms = 148;
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;
	uint32_t W = 0, H = 0;
	unsigned long long* vis = nullptr;
	float* pyramid = nullptr;
	PyramidDesc pyr{};
	uint32_t exact_levels = 0;
	// per-draw buffers (grow-only)
	uint32_t cap_draws = 0;
	uint32_t* list_visible[2] = {nullptr, nullptr};
	uint32_t* list_occluded[2] = {nullptr, nullptr};
	uint32_t* list_tmp = nullptr;
	uint32_t* tmp_count = nullptr;
	uint8_t* status[2] = {nullptr, nullptr};
	bool status_valid[2] = {false, false};
	FrameCounters* counters = nullptr;
	FrameCounters* h_counters = nullptr; // pinned
	// per-transform mvp / determinant sign (mesh.glsl:44,71 hoisted; grow-only)
	float* xf_mvp = nullptr;
	uint32_t* xf_det = nullptr;
	float4* xf_eye = nullptr;         // camera position in each mesh-node's own space (cone cull)
	uint64_t cone_table = 0;          // vkv_set_cone_table
	uint64_t qtable = 0;              // vkv_set_quantized_positions
	uint32_t xf_cap = 0;
	uint32_t xf_count = 0;
	BigTri* big_tris = nullptr;       // large-triangle queue (raster.cu)
	uint32_t big_cap = 1u << 20;      // VKV_BIG_CAP overrides (tests force the overflow path with a tiny queue)
	ClipTri* clip_tris = nullptr;     // clip queue (raster.cu)
	uint32_t clip_cap = 1u << 18;     // VKV_CLIP_CAP overrides
	// multi-GPU: this GPU's shard of the draw list and the peers' visbuffers mapped through CUDA IPC
	uint32_t shard_first = 0, shard_count = 0;       // contiguous shard
	uint32_t shard_block_log2 = 0, shard_rank = 0, shard_nranks = 1; // interleaved shard (blocks of 2^k draws, round-robin)
	bool sharded = false;
	bool no_pdl = false;         // VKV_NO_PDL=1: launch the pass-B cull without programmatic stream serialization (A/B measurements)
	bool separate_clear = false; // VKV_SEPARATE_CLEAR=1: keep the visbuffer clear a launch of its own (A/B measurements)
	// drain launches sized from what the last observed frames found queued (FrameCounters::drain_seen): idle[pass] = consecutive observed
	// frames with empty queues in that pass; the small grid is used from 2 on.  VKV_DRAIN_FULL=1 switches the heuristic off.
	uint32_t drain_idle[2] = {0, 0};
	bool drain_full = false;
	bool full_hiz_b = false;     // VKV_HIZ_FULL_B=1: the pass-B pyramid build redoes every tile, not only the ones pass B drew into (A/B measurements)
	MergeParams mp{};
	bool attached = false;
	// exchange block: ONE allocation the peers map through CUDA IPC — barrier slots (kMaxRanks u32 + 1 error word, first 256 B),
	// the dirty-tile flags of both passes (strip mode), the pyramid.  Rebuilt with the targets (vkv_resize).
	unsigned char* xchg = nullptr;
	size_t xchg_bytes = 0;
	uint32_t* sync_flags = nullptr;   // = xchg
	uint8_t* dirty = nullptr;         // = xchg + 256: [2][dirty_stride]
	uint32_t dirty_stride = 0, tiles_x = 0, tiles_y = 0;
	size_t pyramid_offset = 0;        // of the pyramid inside xchg
	uint32_t epoch = 0;
	bool merge_used = false;          // a barrier kernel has been enqueued since the error word was last read
	bool merge_err_pending = false;
	// resolve pass: RGBA8 target + per-material colour table (grow-only)
	uint32_t* color = nullptr;
	uint32_t* motion = nullptr;       // motion-vector target (R16G16_SFLOAT texels), allocated by the first vkv_motion_vectors
	uint32_t* mat_colors = nullptr;
	uint32_t mat_cap = 0;
	// readback scratch
	uint32_t* tmp_ids = nullptr;
	float* tmp_depth = nullptr;
	void* flush_buf = nullptr;
	size_t flush_bytes = 0;
	cudaEvent_t events[16] = {};
	cudaEvent_t stage_ev[12] = {};
	// frames in flight (vkv_frame_submit / vkv_frame_wait, vkv_update_staged): the reference keeps frameOverlap frames in flight with
	// per-frame camera / draw buffers and fences (application.cpp:133,153,642); here a ring of pinned counter blocks + events
	static constexpr uint32_t kFlights = 4;
	struct Flight { cudaEvent_t done = nullptr; FrameCounters* h = nullptr; uint32_t ticket = 0, draws = 0, launches = 0; bool two = false; };
	Flight flight[kFlights];
	uint32_t next_ticket = 1;
	cudaStream_t upload_stream = nullptr; // staged uploads run beside the previous frame's kernels
	cudaEvent_t upload_ev = nullptr;
	bool upload_pending = false;
	std::map<uint64_t, size_t> allocs;
	std::mutex mtx;
	std::string err;
};

namespace {

thread_local std::string g_create_err;

int fail(vkv_ctx* c, int code, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (c) c->err = buf; else g_create_err = buf;
	return code;
}

#define CK(call)                                                                                                      \
	do {                                                                                                              \
		cudaError_t e_ = (call);                                                                                      \
		if (e_ != cudaSuccess) return fail(c, e_ == cudaErrorMemoryAllocation ? VKV_ERR_OOM : VKV_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
	} while (0)

void fill_pyramid_desc(uint32_t W, uint32_t H, PyramidDesc& d, uint32_t& exact) {
	memset(&d, 0, sizeof(d));
	d.levels = vkv_mip_levels(W, H);
	if (d.levels > 16) d.levels = 16;
	uint32_t off = 0;
	for (uint32_t k = 0; k < d.levels; ++k) {
		d.w[k] = vkv_mip_extent(W, k);
		d.h[k] = vkv_mip_extent(H, k);
		d.off[k] = off;
		off += d.w[k] * d.h[k];
	}
	d.off[d.levels] = off;
	d.total = off;
	// leading mips whose source is exactly 2x their dispatch size (application.cpp:965: levelSize = res >> i)
	exact = 0;
	for (uint32_t k = 0; k < d.levels && k < 4; ++k) {
		const uint32_t sw = k == 0 ? W : d.w[k - 1], sh = k == 0 ? H : d.h[k - 1];
		const uint32_t dw = W >> (k + 1), dh = H >> (k + 1);
		if (dw == 0 || dh == 0 || sw != 2 * dw || sh != 2 * dh || dw != d.w[k] || dh != d.h[k]) break;
		exact = k + 1;
	}
}

int free_targets(vkv_ctx* c) {
	if (c->vis) cudaFree(c->vis);
	if (c->xchg) cudaFree(c->xchg);
	c->xchg = nullptr; c->sync_flags = nullptr; c->dirty = nullptr;
	if (c->tmp_ids) cudaFree(c->tmp_ids);
	if (c->tmp_depth) cudaFree(c->tmp_depth);
	if (c->color) cudaFree(c->color);
	if (c->motion) cudaFree(c->motion);
	c->motion = nullptr;
	c->vis = nullptr; c->pyramid = nullptr; c->tmp_ids = nullptr; c->tmp_depth = nullptr; c->color = nullptr;
	return 0;
}

int alloc_targets(vkv_ctx* c, uint32_t W, uint32_t H) {
	if (W < 2 || H < 2 || W > 32768 || H > 32768) return fail(c, VKV_ERR_INVALID, "resolution %ux%u out of range", W, H);
	free_targets(c);
	c->W = W; c->H = H;
	fill_pyramid_desc(W, H, c->pyr, c->exact_levels);
	CK(cudaMalloc(&c->vis, (size_t)W * H * 8));
	c->tiles_x = (W + 63) / 64; c->tiles_y = (H + 15) / 16;                       // 64x16-pixel tiles (hiz_tile.cuh)
	c->dirty_stride = (c->tiles_x * c->tiles_y + 15u) & ~15u;
	c->pyramid_offset = (256 + 2 * (size_t)c->dirty_stride + 255) & ~(size_t)255;
	c->xchg_bytes = c->pyramid_offset + (((size_t)c->pyr.total + 2) & ~(size_t)1) * 4;   // an even number of floats (vkv_hash reads u64 words)
	CK(cudaMalloc(&c->xchg, c->xchg_bytes));
	c->sync_flags = (uint32_t*)c->xchg;
	c->dirty = c->xchg + 256;
	c->pyramid = (float*)(c->xchg + c->pyramid_offset);
	// initial contents: visbuffer cleared; pyramid 0.0 everywhere = "far" (nothing occludes); barrier slots and flags zero
	CK(launch_fill64(c->vis, (size_t)W * H, VKV_VIS64_CLEAR, c->num_sms, c->stream));
	CK(cudaMemsetAsync(c->xchg, 0, c->xchg_bytes, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return VKV_OK;
}

int ensure_draws(vkv_ctx* c, uint32_t n) {
	if (n > VKV_MAX_MESHLET_DRAWS) return fail(c, VKV_ERR_LIMIT, "meshletDrawCount %u exceeds 2^25 (visbuffer.h.glsl:15-17)", n);
	if (n <= c->cap_draws) return VKV_OK;
	CK(cudaStreamSynchronize(c->stream));
	const uint32_t cap = n + n / 8 + 1024;
	// allocate the whole new set first and swap it in only when every allocation succeeded: a failure in the middle leaves the
	// context exactly as it was (old buffers, old capacity)
	void* fresh[7] = {};
	const size_t bytes[7] = {(size_t)cap * 4, (size_t)cap * 4, (size_t)cap * 4, (size_t)cap * 4, (size_t)cap, (size_t)cap, (size_t)cap * 4};
	for (int i = 0; i < 7; ++i) {
		cudaError_t e = cudaMalloc(&fresh[i], bytes[i]);
		if (e != cudaSuccess) {
			for (int k = 0; k < i; ++k) cudaFree(fresh[k]);
			return fail(c, e == cudaErrorMemoryAllocation ? VKV_ERR_OOM : VKV_ERR_CUDA, "growing the per-draw buffers to %u draws: %s", cap, cudaGetErrorString(e));
		}
	}
	for (int i = 0; i < 2; ++i) {
		if (c->list_visible[i]) cudaFree(c->list_visible[i]);
		if (c->list_occluded[i]) cudaFree(c->list_occluded[i]);
		if (c->status[i]) cudaFree(c->status[i]);
		c->list_visible[i] = (uint32_t*)fresh[i];
		c->list_occluded[i] = (uint32_t*)fresh[2 + i];
		c->status[i] = (uint8_t*)fresh[4 + i];
		c->status_valid[i] = false;
	}
	if (c->list_tmp) cudaFree(c->list_tmp);
	c->list_tmp = (uint32_t*)fresh[6];
	c->cap_draws = cap;
	return VKV_OK;
}

// mesh.glsl:43-44,71 once per mesh-node: needs the transform count, which the push constants do not carry — it is the
// extent of the vkv_upload allocation the transform buffer lives in.
// `fused` != NULL: only size the buffers and hand the job to the pass-A cull launch (cull.cu) through its parameters
int prepare_transforms(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, int* launches, CullParams* fused = nullptr, bool want_eye = false) {
	c->xf_count = 0;
	if (!pc->meshletDrawCount) return VKV_OK;
	size_t bytes = 0;
	{
		std::lock_guard<std::mutex> lock(c->mtx);
		auto it = c->allocs.upper_bound(pc->transformBuffer);
		if (it != c->allocs.begin()) {
			--it;
			if (pc->transformBuffer < it->first + it->second) bytes = it->first + it->second - pc->transformBuffer;
		}
	}
	if (bytes < 64) return fail(c, VKV_ERR_INVALID, "transformBuffer does not point into a vkv_upload allocation");
	const size_t n64 = bytes / 64;
	if (n64 > 0xffffffffull) return fail(c, VKV_ERR_INVALID, "transform buffer too large");
	const uint32_t n = (uint32_t)n64;
	if (n > c->xf_cap) {
		CK(cudaStreamSynchronize(c->stream));
		const uint32_t cap = n + n / 8 + 64;
		float* mvp = nullptr; uint32_t* det = nullptr; float4* eye = nullptr;
		CK(cudaMalloc(&mvp, (size_t)cap * 64));
		cudaError_t e2 = cudaMalloc(&det, (size_t)cap * 4);
		if (e2 == cudaSuccess) e2 = cudaMalloc(&eye, (size_t)cap * 16);
		if (e2 != cudaSuccess) { cudaFree(mvp); if (det) cudaFree(det); return fail(c, VKV_ERR_OOM, "growing the per-transform buffers: %s", cudaGetErrorString(e2)); }
		if (c->xf_mvp) cudaFree(c->xf_mvp);
		if (c->xf_det) cudaFree(c->xf_det);
		if (c->xf_eye) cudaFree(c->xf_eye);
		c->xf_mvp = mvp; c->xf_det = det; c->xf_eye = eye; c->xf_cap = cap;
	}
	if (fused) { fused->xf_mvp = c->xf_mvp; fused->xf_det = c->xf_det; fused->xf_n = n; }
	else {
		CK(launch_prepare_transforms((const float*)pc->transformBuffer, (const vkv_Camera*)pc->cameraBuffer, n, c->xf_mvp, c->xf_det, want_eye ? c->xf_eye : nullptr,
		                             c->num_sms, c->stream));
		if (launches) ++*launches;
	}
	c->xf_count = n;
	return VKV_OK;
}

CullParams make_cull(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, int pass, uint32_t flags) {
	CullParams p{};
	p.draws = (const vkv_MeshletDraw*)pc->drawBuffer;
	p.transforms = (const float*)pc->transformBuffer;
	p.primitives = (const vkv_Primitive*)pc->primitiveBuffer;
	p.camera = (const vkv_Camera*)pc->cameraBuffer;
	p.pyramid = c->pyramid;
	p.pyr = c->pyr;
	p.total = pc->meshletDrawCount;
	p.n = pc->meshletDrawCount;
	if (c->sharded && c->shard_block_log2) { // local index space: my blocks, the last one possibly ragged (bounded by `total` in the kernel)
		const uint32_t B = c->shard_block_log2;
		const uint64_t blocks = ((uint64_t)pc->meshletDrawCount + (1u << B) - 1) >> B;
		const uint64_t mine = blocks > c->shard_rank ? (blocks - c->shard_rank + c->shard_nranks - 1) / c->shard_nranks : 0;
		p.n = (uint32_t)(mine << B);
		p.shard_block_log2 = B; p.shard_rank = c->shard_rank; p.shard_nranks = c->shard_nranks;
	} else if (c->sharded) {
		p.n = c->shard_count;
		p.first = c->shard_first;
	}
	p.in_list = pass == 0 ? nullptr : c->list_occluded[0];
	p.in_count = pass == 0 ? nullptr : &c->counters->occluded[0];
	p.out_visible = c->list_visible[pass];
	p.out_occluded = c->list_occluded[pass];
	p.counters = c->counters;
	p.status = (flags & VKV_FRAME_STATUS) ? c->status[pass] : nullptr;
	if ((flags & VKV_FRAME_CONE_CULL) && pass == 0) { p.cone_table = (const unsigned long long*)(uintptr_t)c->cone_table; p.xf_eye = c->xf_eye; }
	p.pass = pass;
	p.vp_select = pass;
	p.skip_hiz = 0;
	p.neg_zero2 = kNegZero2;
	return p;
}

RasterParams make_raster(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, const uint32_t* list, const uint32_t* count, uint32_t* work, int mark_pass = -1) {
	RasterParams r{};
	if (mark_pass >= 0) { r.dirty = c->dirty + (size_t)mark_pass * c->dirty_stride; r.dirtyTilesX = c->tiles_x; r.markLimit = 0xffffffffu; } // strip mode (vkv_frame lowers markLimit for the pass-B pyramid rebuild)
	r.draws = (const vkv_MeshletDraw*)pc->drawBuffer;
	r.transforms = (const float*)pc->transformBuffer;
	r.primitives = (const vkv_Primitive*)pc->primitiveBuffer;
	r.materials = (const vkv_Material*)pc->materialBuffer;
	r.camera = (const vkv_Camera*)pc->cameraBuffer;
	r.list = list; r.count = count; r.work = work;
	r.vis = c->vis; r.W = c->W; r.H = c->H;
	r.neg_zero2 = kNegZero2;
	r.mvp = c->xf_mvp; r.detNeg = c->xf_det;
	r.big = c->big_tris; r.bigCap = c->big_cap; r.bigCursor = &c->counters->big_cursor; r.bigNext = &c->counters->big_next;
	r.clip = c->clip_tris; r.clipCap = c->clip_cap; r.clipCount = &c->counters->clip_count; r.clipNext = &c->counters->clip_next;
	r.qtable = (const vkv_QuantizedPositions*)(uintptr_t)c->qtable;
	r.overflow = &c->counters->raster_overflow; r.drainBarrier = &c->counters->drain_barrier; r.slowWork = &c->counters->slow_work;
	return r;
}

// raster_kernel + raster_big_kernel behind a reset of the large-triangle queue
// after_cull: the cull launch right before this call has zeroed the raster's counters itself (CullParams::reset_ptr) and nothing
// sits between the two launches: the raster kernel is launched with programmatic stream serialization
int enqueue_raster(vkv_ctx* c, RasterParams r, int* launches, bool after_cull = false, int pass = -1, bool multi_gpu = false) {
	if (!after_cull) CK(cudaMemsetAsync(&c->counters->big_next, 0, offsetof(FrameCounters, raster_reset_end) - offsetof(FrameCounters, big_next), c->stream)); // queues, cursors, barrier
	bool small_drain = false;
	// (single-GPU frames only: the multi-GPU exchange paths keep the launch shape they were validated with)
	if (pass >= 0) { r.drainSeen = &c->counters->drain_seen[pass]; small_drain = !c->drain_full && !multi_gpu && c->drain_idle[pass] >= 2; }
	CK(launch_raster(r, c->num_sms, c->stream, after_cull, small_drain));
	if (launches) *launches += 2;
	return VKV_OK;
}

HizParams make_hiz(vkv_ctx* c) {
	HizParams h{};
	h.vis = c->vis; h.pyramid = c->pyramid; h.pyr = c->pyr; h.W = c->W; h.H = c->H; h.exact_levels = c->exact_levels; h.done = (getenv("VKV_HIZ_NOTAIL") || getenv("VKV_HIZ_SPLIT")) ? nullptr : &c->counters->hiz_done; /* diagnosis switches */
	h.split_tail = getenv("VKV_HIZ_SPLIT") ? 1 : 0;
	return h;
}

// what a frame's drain kernels found queued -> how the next frames' drain launches are sized (enqueue_raster)
void observe_drain(vkv_ctx* c, const FrameCounters* h, bool two) {
	for (int pass = 0; pass < (two ? 2 : 1); ++pass) c->drain_idle[pass] = h->drain_seen[pass] ? 0u : c->drain_idle[pass] + 1u;
}

// staged uploads (vkv_update_staged) become visible to everything enqueued on the context's stream from here on
// Pass-B pyramid rebuild: the rasteriser marks tiles (and the pyramid kernel skips clean ones) only while pass B draws at most this many
// meshlets.  Measured (profiles/r4e, r4f): 7 k meshlets (cfg 3): pyramid B 29.1 -> 24 us for ~1-4 us of marks; 54 k (cfg 4): marks cost
// 5-15 us and every tile is dirty anyway.
constexpr uint32_t kPartialHizLimit = 16384;

int join_uploads(vkv_ctx* c) {
	if (!c->upload_pending) return VKV_OK;
	CK(cudaEventRecord(c->upload_ev, c->upload_stream));
	CK(cudaStreamWaitEvent(c->stream, c->upload_ev, 0));
	c->upload_pending = false;
	return VKV_OK;
}

int check_pc(vkv_ctx* c, const vkv_VisbufferPushConstants* pc) {
	if (!c) return VKV_ERR_INVALID;
	if (!pc) return fail(c, VKV_ERR_INVALID, "push constants are NULL");
	if (pc->meshletDrawCount && (!pc->drawBuffer || !pc->transformBuffer || !pc->primitiveBuffer || !pc->cameraBuffer || !pc->materialBuffer))
		return fail(c, VKV_ERR_INVALID, "push constants hold a NULL buffer address");
	if (c->sharded && !c->shard_block_log2 && (uint64_t)c->shard_first + c->shard_count > pc->meshletDrawCount)
		return fail(c, VKV_ERR_INVALID, "draw shard [%u, +%u) exceeds meshletDrawCount %u", c->shard_first, c->shard_count, pc->meshletDrawCount);
	return ensure_draws(c, pc->meshletDrawCount);
}

void detach_peers(vkv_ctx* c) {
	if (!c->attached) return;
	for (int r = 0; r < c->mp.nranks; ++r) {
		if (r == c->mp.rank) continue;
		if (c->mp.vis[r]) cudaIpcCloseMemHandle(c->mp.vis[r]);
		if (c->mp.flags[r]) cudaIpcCloseMemHandle(c->mp.flags[r]); // the peer's exchange block (flags, dirty, pyramid share one mapping)
	}
	memset(&c->mp, 0, sizeof(c->mp));
	c->attached = false;
}

// barrier -> fused reduce-scatter/all-gather min over peer memory -> barrier (merge.cu)
int enqueue_merge(vkv_ctx* c, int* launches) {
	if (!c->attached) return fail(c, VKV_ERR_INVALID, "merge requested but no peers are attached (vkv_ipc_attach)");
	const unsigned long long timeout_ns = 5ull * 1000 * 1000 * 1000;
	c->mp.n = (size_t)c->W * c->H;
	CK(launch_xgpu_barrier(c->mp, ++c->epoch, timeout_ns, c->stream));
	CK(launch_merge_min(c->mp, c->num_sms, c->stream));
	CK(launch_xgpu_barrier(c->mp, ++c->epoch, timeout_ns, c->stream));
	c->merge_used = true;
	if (launches) *launches += 3;
	return VKV_OK;
}

// strip mode (strips.cu): barrier -> [pull dirty tiles of my strip, min, exact mips, changed texels to every pyramid] -> barrier ->
// small mips locally
int enqueue_strip_exchange(vkv_ctx* c, int pass, int* launches) {
	if (!c->attached) return fail(c, VKV_ERR_INVALID, "strip exchange requested but no peers are attached (vkv_ipc_attach)");
	const unsigned long long timeout_ns = 5ull * 1000 * 1000 * 1000;
	StripParams sp{};
	sp.mp = c->mp; sp.mp.n = (size_t)c->W * c->H;
	sp.W = c->W; sp.H = c->H; sp.pyr = c->pyr; sp.exact_levels = c->exact_levels;
	sp.tilesX = c->tiles_x; sp.tilesY = c->tiles_y; sp.dirtyStride = c->dirty_stride; sp.pass = pass;
	sp.stats = &c->counters->strip_tiles_pulled;
	// two launches per exchange: the cross-GPU barriers ride inside them (strips.cu, hiz.cu)
	sp.epoch_in = ++c->epoch; sp.epoch_out = ++c->epoch;
	sp.done = &c->counters->strip_done; sp.timeout_ns = timeout_ns;
	static const bool timing = getenv("VKV_STRIP_TIMING") != nullptr; // diagnosis: events between the launches, printed by rank 0
	if (timing) cudaEventRecord(c->events[8], c->stream);
	CK(launch_strip_merge_hiz(sp, c->num_sms, c->stream));
	if (timing) cudaEventRecord(c->events[9], c->stream);
	HizParams h = make_hiz(c);
	h.wait_flags = c->sync_flags; h.wait_epoch = sp.epoch_out; h.wait_ranks = c->mp.nranks; h.wait_error = c->sync_flags + kMaxRanks; h.wait_timeout_ns = timeout_ns;
	if (c->exact_levels < c->pyr.levels) CK(launch_hiz_tail(h, c->stream));
	else CK(launch_xgpu_barrier(c->mp, sp.epoch_out, timeout_ns, c->stream)); // no small mips (tiny targets): the out-barrier needs a launch of its own
	if (timing) {
		cudaEventRecord(c->events[10], c->stream);
		cudaEventSynchronize(c->events[10]);
		float t[2];
		for (int i = 0; i < 2; ++i) cudaEventElapsedTime(&t[i], c->events[8 + i], c->events[9 + i]);
		if (c->mp.rank == 0) fprintf(stderr, "[strip pass %d] barrier + merge + mips %.1f us   barrier + tail %.1f us\n", pass, t[0] * 1e3f, t[1] * 1e3f);
	}
	c->merge_used = true;
	if (launches) *launches += 2;
	return VKV_OK;
}

// The merge barriers report a peer that never arrived through an error word on the device.  Every call that synchronises the
// stream afterwards (frame with stats, vkv_sync, the vkv_read_* family) reads it, fails once with VKV_ERR_CUDA and clears it.
int check_merge_error(vkv_ctx* c) {
	if (!c->merge_used || !c->sync_flags) return VKV_OK;
	uint32_t err = 0;
	CK(cudaMemcpyAsync(&err, c->sync_flags + kMaxRanks, 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	c->merge_used = false;
	if (err) {
		CK(cudaMemsetAsync(c->sync_flags + kMaxRanks, 0, 4, c->stream));
		return fail(c, VKV_ERR_CUDA, "multi-GPU merge barrier timed out (a peer did not reach the merge); the visbuffer of that frame is unmerged");
	}
	return VKV_OK;
}

} // namespace

extern "C" {

const char* vkv_last_error(vkv_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int vkv_create(vkv_ctx** out, int cuda_device, uint32_t width, uint32_t height) {
	if (!out) return VKV_ERR_INVALID;
	*out = nullptr;
	vkv_ctx* c = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return fail(nullptr, VKV_ERR_NO_DEVICE, "no CUDA device (%s); libvkv has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
	if (cuda_device < 0 || cuda_device >= ndev) return fail(nullptr, VKV_ERR_INVALID, "cuda_device %d out of range [0,%d)", cuda_device, ndev);
	c = new vkv_ctx();
	c->device = cuda_device;
	auto bail = [&](int rc) { g_create_err = c->err; vkv_destroy(c); return rc; };
	if (cudaSetDevice(cuda_device) != cudaSuccess) { c->err = "cudaSetDevice failed"; return bail(VKV_ERR_CUDA); }
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, cuda_device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
	if (const char* e = getenv("VKV_SEPARATE_CLEAR")) c->separate_clear = e[0] == '1';
	if (const char* e = getenv("VKV_NO_PDL")) c->no_pdl = e[0] == '1';
	if (const char* e = getenv("VKV_HIZ_FULL_B")) c->full_hiz_b = e[0] == '1';
	if (const char* e = getenv("VKV_DRAIN_FULL")) c->drain_full = e[0] == '1';
	if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { c->err = "cudaStreamCreate failed"; return bail(VKV_ERR_CUDA); }
	c->stream = c->own_stream;
	for (auto& ev : c->events) cudaEventCreate(&ev);
	for (auto& ev : c->stage_ev) cudaEventCreate(&ev);
	if (cudaMalloc(&c->counters, sizeof(FrameCounters)) != cudaSuccess || cudaMalloc(&c->tmp_count, 256) != cudaSuccess ||
	    cudaMallocHost(&c->h_counters, sizeof(FrameCounters)) != cudaSuccess) {
		c->err = "allocating counters failed";
		return bail(VKV_ERR_OOM);
	}
	cudaMemset(c->counters, 0, sizeof(FrameCounters));
	for (auto& f : c->flight)
		if (cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming) != cudaSuccess || cudaMallocHost(&f.h, sizeof(FrameCounters)) != cudaSuccess) {
			c->err = "allocating the frames-in-flight ring failed";
			return bail(VKV_ERR_OOM);
		}
	if (cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreateWithFlags(&c->upload_ev, cudaEventDisableTiming) != cudaSuccess) { c->err = "creating the upload stream failed"; return bail(VKV_ERR_CUDA); }
	if (const char* e = getenv("VKV_BIG_CAP")) c->big_cap = (uint32_t)std::max(1L, std::min(1L << 22, atol(e)));
	if (const char* e = getenv("VKV_CLIP_CAP")) c->clip_cap = (uint32_t)std::max(1L, std::min(1L << 22, atol(e)));
	if (cudaMalloc(&c->big_tris, (size_t)c->big_cap * sizeof(BigTri)) != cudaSuccess || cudaMalloc(&c->clip_tris, (size_t)c->clip_cap * sizeof(ClipTri)) != cudaSuccess) {
		c->err = "allocating the large-triangle / clip queues failed";
		return bail(VKV_ERR_OOM);
	}
	int rc = alloc_targets(c, width, height);
	if (rc != VKV_OK) return bail(rc);
	*out = c;
	return VKV_OK;
}

int vkv_resize(vkv_ctx* c, uint32_t width, uint32_t height) {
	if (!c) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	detach_peers(c); // the visbuffer is reallocated: peers must re-export / re-attach
	return alloc_targets(c, width, height);
}

void vkv_destroy(vkv_ctx* c) {
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	detach_peers(c);
	free_targets(c);
	for (int i = 0; i < 2; ++i) {
		if (c->list_visible[i]) cudaFree(c->list_visible[i]);
		if (c->list_occluded[i]) cudaFree(c->list_occluded[i]);
		if (c->status[i]) cudaFree(c->status[i]);
	}
	if (c->list_tmp) cudaFree(c->list_tmp);
	if (c->xf_mvp) cudaFree(c->xf_mvp);
	if (c->xf_det) cudaFree(c->xf_det);
	if (c->xf_eye) cudaFree(c->xf_eye);
	if (c->big_tris) cudaFree(c->big_tris);
	if (c->clip_tris) cudaFree(c->clip_tris);
	if (c->mat_colors) cudaFree(c->mat_colors);
	if (c->tmp_count) cudaFree(c->tmp_count);
	if (c->counters) cudaFree(c->counters);
	if (c->h_counters) cudaFreeHost(c->h_counters);
	if (c->upload_stream) { cudaStreamSynchronize(c->upload_stream); cudaStreamDestroy(c->upload_stream); }
	if (c->upload_ev) cudaEventDestroy(c->upload_ev);
	for (auto& f : c->flight) { if (f.done) cudaEventDestroy(f.done); if (f.h) cudaFreeHost(f.h); }
	if (c->flush_buf) cudaFree(c->flush_buf);
	for (auto& kv : c->allocs) cudaFree((void*)(uintptr_t)kv.first);
	for (auto& ev : c->events) if (ev) cudaEventDestroy(ev);
	for (auto& ev : c->stage_ev) if (ev) cudaEventDestroy(ev);
	if (c->own_stream) cudaStreamDestroy(c->own_stream);
	delete c;
}

int vkv_set_stream(vkv_ctx* c, void* s) {
	if (!c) return VKV_ERR_INVALID;
	CK(cudaStreamSynchronize(c->stream));
	c->stream = s ? (cudaStream_t)s : c->own_stream;
	return VKV_OK;
}

int vkv_sync(vkv_ctx* c) {
	if (!c) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	{ int rc = join_uploads(c); if (rc) return rc; }
	CK(cudaStreamSynchronize(c->stream));
	return check_merge_error(c);
}

int vkv_upload(vkv_ctx* c, const void* host, size_t bytes, uint64_t* dev_addr) {
	if (!c || !dev_addr || (!host && bytes)) return c ? fail(c, VKV_ERR_INVALID, "vkv_upload: NULL argument") : VKV_ERR_INVALID;
	std::lock_guard<std::mutex> lock(c->mtx);
	CK(cudaSetDevice(c->device));
	void* d = nullptr;
	// pad so that 16-byte vector loads over the tail of any array stay inside the allocation
	const size_t padded = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
	CK(cudaMalloc(&d, padded));
	if (bytes) CK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
	c->allocs[(uint64_t)(uintptr_t)d] = bytes;
	*dev_addr = (uint64_t)(uintptr_t)d;
	return VKV_OK;
}

int vkv_update(vkv_ctx* c, uint64_t dev_addr, const void* host, size_t bytes) {
	if (!c || !host) return VKV_ERR_INVALID;
	{
		std::lock_guard<std::mutex> lock(c->mtx);
		auto it = c->allocs.upper_bound(dev_addr);
		if (it == c->allocs.begin()) return fail(c, VKV_ERR_INVALID, "vkv_update: unknown address");
		--it;
		if (dev_addr + bytes > it->first + it->second) return fail(c, VKV_ERR_INVALID, "vkv_update: range exceeds the allocation");
	}
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync((void*)(uintptr_t)dev_addr, host, bytes, cudaMemcpyHostToDevice, c->stream));
	return VKV_OK;
}

int vkv_update_staged(vkv_ctx* c, uint64_t dev_addr, const void* pinned_host, size_t bytes) {
	if (!c || !pinned_host) return VKV_ERR_INVALID;
	{
		std::lock_guard<std::mutex> lock(c->mtx);
		auto it = c->allocs.upper_bound(dev_addr);
		if (it == c->allocs.begin()) return fail(c, VKV_ERR_INVALID, "vkv_update_staged: unknown address");
		--it;
		if (dev_addr + bytes > it->first + it->second) return fail(c, VKV_ERR_INVALID, "vkv_update_staged: range exceeds the allocation");
	}
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync((void*)(uintptr_t)dev_addr, pinned_host, bytes, cudaMemcpyHostToDevice, c->upload_stream));
	c->upload_pending = true; // the next frame / stage call on the context's stream waits for it (join_uploads)
	return VKV_OK;
}

int vkv_free(vkv_ctx* c, uint64_t dev_addr) {
	if (!c) return VKV_ERR_INVALID;
	std::lock_guard<std::mutex> lock(c->mtx);
	auto it = c->allocs.find(dev_addr);
	if (it == c->allocs.end()) return fail(c, VKV_ERR_INVALID, "vkv_free: unknown address");
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaFree((void*)(uintptr_t)dev_addr));
	c->allocs.erase(it);
	return VKV_OK;
}

int vkv_set_cone_table(vkv_ctx* c, uint64_t table_dev_addr) {
	if (!c) return VKV_ERR_INVALID;
	if (table_dev_addr) {
		std::lock_guard<std::mutex> lock(c->mtx);
		auto it = c->allocs.upper_bound(table_dev_addr);
		if (it == c->allocs.begin() || table_dev_addr >= std::prev(it)->first + std::prev(it)->second)
			return fail(c, VKV_ERR_INVALID, "vkv_set_cone_table: the table does not lie in a vkv_upload allocation");
	}
	c->cone_table = table_dev_addr;
	return VKV_OK;
}

int vkv_set_quantized_positions(vkv_ctx* c, uint64_t table_dev_addr) {
	if (!c) return VKV_ERR_INVALID;
	if (table_dev_addr) {
		std::lock_guard<std::mutex> lock(c->mtx);
		auto it = c->allocs.upper_bound(table_dev_addr);
		if (it == c->allocs.begin() || table_dev_addr >= std::prev(it)->first + std::prev(it)->second)
			return fail(c, VKV_ERR_INVALID, "vkv_set_quantized_positions: the table does not lie in a vkv_upload allocation");
	}
	c->qtable = table_dev_addr;
	return VKV_OK;
}

int vkv_clear(vkv_ctx* c) {
	if (!c) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	CK(launch_fill64(c->vis, (size_t)c->W * c->H, VKV_VIS64_CLEAR, c->num_sms, c->stream)); // application.cpp:782,807
	return VKV_OK;
}

int vkv_cull(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, int pass, uint32_t flags, uint32_t* n_visible) {
	int rc = check_pc(c, pc);
	if (rc) return rc;
	if (pass < 0 || pass > 1) return fail(c, VKV_ERR_INVALID, "pass must be 0 or 1");
	CK(cudaSetDevice(c->device));
	CK(cudaMemsetAsync(&c->counters->visible[pass], 0, 4, c->stream));
	CK(cudaMemsetAsync(&c->counters->occluded[pass], 0, 4, c->stream));
	if ((flags & VKV_FRAME_CONE_CULL) && pass == 0) {
		if (!c->cone_table) return fail(c, VKV_ERR_INVALID, "VKV_FRAME_CONE_CULL needs vkv_set_cone_table first");
		rc = prepare_transforms(c, pc, nullptr, nullptr, true);
		if (rc) return rc;
	}
	CullParams p = make_cull(c, pc, pass, flags);
	if (p.status) CK(cudaMemsetAsync(p.status, VKV_ST_NOT_TESTED, pc->meshletDrawCount, c->stream));
	c->status_valid[pass] = p.status != nullptr;
	if (p.n) CK(launch_cull(p, c->num_sms, c->stream));
	if (n_visible) {
		CK(cudaMemcpyAsync(c->h_counters, c->counters, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		*n_visible = c->h_counters->visible[pass];
	}
	return VKV_OK;
}

int vkv_raster(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, int pass) {
	int rc = check_pc(c, pc);
	if (rc) return rc;
	if (pass < 0 || pass > 1) return fail(c, VKV_ERR_INVALID, "pass must be 0 or 1");
	CK(cudaSetDevice(c->device));
	CK(cudaMemsetAsync(&c->counters->work[pass], 0, 4, c->stream));
	rc = prepare_transforms(c, pc, nullptr);
	if (rc) return rc;
	rc = enqueue_raster(c, make_raster(c, pc, c->list_visible[pass], &c->counters->visible[pass], &c->counters->work[pass]), nullptr);
	if (rc) return rc;
	return VKV_OK;
}

int vkv_raster_list(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, const uint32_t* draw_ids, uint32_t n) {
	int rc = check_pc(c, pc);
	if (rc) return rc;
	if (n > c->cap_draws) { rc = ensure_draws(c, n); if (rc) return rc; }
	CK(cudaSetDevice(c->device));
	uint32_t hdr[2] = {n, 0};
	CK(cudaMemcpyAsync(c->tmp_count, hdr, 8, cudaMemcpyHostToDevice, c->stream));
	if (n) CK(cudaMemcpyAsync(c->list_tmp, draw_ids, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
	rc = prepare_transforms(c, pc, nullptr);
	if (rc) return rc;
	rc = enqueue_raster(c, make_raster(c, pc, c->list_tmp, c->tmp_count, c->tmp_count + 1), nullptr);
	if (rc) return rc;
	CK(cudaStreamSynchronize(c->stream)); // draw_ids / hdr are borrowed
	return VKV_OK;
}

int vkv_hiz(vkv_ctx* c) {
	if (!c) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	CK(launch_hiz(make_hiz(c), c->num_sms, c->stream, nullptr));
	return VKV_OK;
}

namespace {
int frame_impl(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, uint32_t flags, vkv_stats* out, uint32_t* launches_out) {
	int rc = check_pc(c, pc);
	if (rc) return rc;
	CK(cudaSetDevice(c->device));
	rc = join_uploads(c);
	if (rc) return rc;
	cudaStream_t s = c->stream;
	const bool timed = (flags & VKV_FRAME_TIMED) && out;
	const bool two = (flags & VKV_FRAME_TWO_PASS) && !(flags & VKV_FRAME_NO_CULL);
	const bool strips = (flags & VKV_FRAME_MERGE_STRIPS) != 0;
	const bool merge = (flags & VKV_FRAME_MERGE) != 0 && !strips;
	const bool hiz = !(flags & VKV_FRAME_NO_HIZ);
	const uint32_t N = pc->meshletDrawCount;
	if ((merge || strips) && !c->attached) return fail(c, VKV_ERR_INVALID, "VKV_FRAME_MERGE / VKV_FRAME_MERGE_STRIPS need vkv_ipc_attach first");
	const bool cone = (flags & VKV_FRAME_CONE_CULL) && !(flags & VKV_FRAME_NO_CULL);
	if (cone && !c->cone_table) return fail(c, VKV_ERR_INVALID, "VKV_FRAME_CONE_CULL needs vkv_set_cone_table first");
	if (strips && (c->exact_levels < 1 || !hiz)) return fail(c, VKV_ERR_INVALID, "VKV_FRAME_MERGE_STRIPS needs an even resolution (one exact pyramid mip) and the pyramid rebuild");
	int launches = 0;
	// NVTX ranges named like the reference's Tracy/debug-label zones (application.cpp:765 "Visbuffer pass", :952 "HiZ reduction") so a
	// maintainer can line a trace of this library up with a trace of the Vulkan path; free when no tool is attached
	struct Range { // closes itself on every early return
		bool open = true;
		explicit Range(const char* n) { nvtxRangePushA(n); }
		void end() { if (open) { nvtxRangePop(); open = false; } }
		~Range() { end(); }
	};
	enum { E_BEGIN, E_CLEAR, E_CULL_A, E_RASTER_A, E_MERGE_A, E_HIZ_A, E_CULL_B, E_RASTER_B, E_MERGE_B, E_HIZ_B, E_COUNT };
	const bool stages = timed && (flags & VKV_FRAME_STAGES);
	auto mark = [&](int e) { if (stages || (timed && e == E_BEGIN)) cudaEventRecord(c->stage_ev[e], s); };
	mark(E_BEGIN); // the counter reset below is a mandatory operation of every frame: inside the timed region
	CK(cudaMemsetAsync(c->counters, 0, sizeof(FrameCounters), s));
	// application.cpp:782,807 — the clear rides inside the pass-A cull launch (cull.cu) whenever there is one and the pixel
	// count is even (16-byte stores); clear_ms then reads ~0 and cull_a_ms covers both
	const size_t npix = (size_t)c->W * c->H;
	bool xf_done = false;
	Range visA("Visbuffer pass");
	CullParams pa = make_cull(c, pc, 0, (flags & VKV_FRAME_NO_CULL) ? 0 : flags);
	const bool fuse_clear = !(flags & VKV_FRAME_NO_CULL) && pa.n > 0 && (npix & 1) == 0 && !c->separate_clear;
	if (!fuse_clear) { CK(launch_fill64(c->vis, npix, VKV_VIS64_CLEAR, c->num_sms, s)); ++launches; }
	mark(E_CLEAR);
	if (flags & VKV_FRAME_NO_CULL) {
		const CullParams& p = pa;
		CK(launch_iota(p, c->list_visible[0], &c->counters->visible[0], c->num_sms, s)); ++launches;
		c->status_valid[0] = false;
	} else {
		CullParams& p = pa;
		if (fuse_clear) { p.clear_ptr = (ulonglong2*)c->vis; p.clear_n2 = npix / 2; p.clear_value = VKV_VIS64_CLEAR; }
		if (strips && p.n) { p.zero_ptr = (uint4*)c->dirty; p.zero_n16 = 2 * c->dirty_stride / 16; } // the dirty-tile flags ride along too
		if (p.status) CK(cudaMemsetAsync(p.status, VKV_ST_NOT_TESTED, N, s));
		c->status_valid[0] = p.status != nullptr;
		if (p.n && !stages) { // (a stage event between cull and raster would separate the two launches anyway)
			p.reset_ptr = &c->counters->big_next;
			p.reset_words = (uint32_t)((offsetof(FrameCounters, raster_reset_end) - offsetof(FrameCounters, big_next)) / 4);
		}
		if (p.n) {
			// the cone test reads the per-node eye positions in THIS launch: they need a launch of their own ahead of it
			if (cone) { rc = prepare_transforms(c, pc, &launches, nullptr, true); if (rc) return rc; xf_done = true; p.xf_eye = c->xf_eye; }
			else if (!c->separate_clear) { rc = prepare_transforms(c, pc, &launches, &p); if (rc) return rc; xf_done = true; } // rides along, like the clear
			CK(launch_cull(p, c->num_sms, s)); ++launches;
		}
	}
	mark(E_CULL_A);
	if (strips && !pa.zero_ptr) CK(cudaMemsetAsync(c->dirty, 0, 2 * (size_t)c->dirty_stride, s));
	// cull A -> raster A back to back (no transform launch, no dirty-flag memset in between): the cull zeroed the raster's counters
	const bool chainA = pa.reset_ptr != nullptr && xf_done && !(strips && !pa.zero_ptr);
	if (!xf_done) { rc = prepare_transforms(c, pc, &launches); if (rc) return rc; }
	rc = enqueue_raster(c, make_raster(c, pc, c->list_visible[0], &c->counters->visible[0], &c->counters->work[0], strips ? 0 : -1), &launches, chainA, 0, merge || strips);
	if (rc) return rc;
	mark(E_RASTER_A);
	if (merge) { rc = enqueue_merge(c, &launches); if (rc) return rc; }
	if (strips) { rc = enqueue_strip_exchange(c, 0, &launches); if (rc) return rc; } // merge + exact mips + all-gather + small mips
	mark(E_MERGE_A);
	visA.end();
	if (hiz && !strips) { Range z("HiZ reduction"); CK(launch_hiz(make_hiz(c), c->num_sms, s, &launches)); }
	mark(E_HIZ_A);
	if (two) {
		Range visB("Visbuffer pass (B)");
		CullParams p = make_cull(c, pc, 1, flags);
		p.skip_frustum = 1; // same camera buffer, same frustum planes as pass A a few launches ago: its survivors pass again
		if (p.status) CK(cudaMemsetAsync(p.status, VKV_ST_NOT_TESTED, N, s));
		c->status_valid[1] = p.status != nullptr;
		// nothing between the pyramid launch and this one (no stage event, no status memset, no merge): let it start under the tail
		const bool pdl = hiz && !stages && !p.status && c->exact_levels >= 1 && !c->no_pdl && !strips;
		const bool chainB = p.n && !stages;
		if (chainB) {
			p.reset_ptr = &c->counters->big_next;
			p.reset_words = (uint32_t)((offsetof(FrameCounters, raster_reset_end) - offsetof(FrameCounters, big_next)) / 4);
		}
		// The second pyramid build only has to redo what pass B changes: the pass-B rasteriser marks every 64x16-pixel tile it may draw
		// into (the dirty bytes of strip mode, set B), the tiled kernel skips the others — their exact mips were stored by the pass-A
		// build a moment ago — and the small-mip tail runs as always.  The flags are zeroed by the pass-B cull launch.
		const bool partialB = hiz && !strips && !merge && c->exact_levels >= 1 && !c->full_hiz_b;
		uint8_t* const dirtyB = c->dirty + c->dirty_stride;
		if (partialB) {
			if (p.n) { p.zero_ptr = (uint4*)dirtyB; p.zero_n16 = c->dirty_stride / 16; }
			else CK(cudaMemsetAsync(dirtyB, 0, c->dirty_stride, s));
		}
		if (p.n) { CK(launch_cull(p, c->num_sms, s, pdl)); ++launches; }
		mark(E_CULL_B);
		RasterParams rb = make_raster(c, pc, c->list_visible[1], &c->counters->visible[1], &c->counters->work[1], (strips || partialB) ? 1 : -1);
		if (partialB) rb.markLimit = kPartialHizLimit;
		rc = enqueue_raster(c, rb, &launches, chainB, 1, merge || strips);
		if (rc) return rc;
		mark(E_RASTER_B);
		if (merge) { rc = enqueue_merge(c, &launches); if (rc) return rc; }
		if (strips) { rc = enqueue_strip_exchange(c, 1, &launches); if (rc) return rc; }
		mark(E_MERGE_B);
		visB.end();
		if (hiz && !strips) {
			Range z("HiZ reduction");
			HizParams hp = make_hiz(c);
			if (partialB) { hp.tile_dirty = dirtyB; hp.dirty_count = &c->counters->visible[1]; hp.dirty_limit = kPartialHizLimit; }
			if (c->exact_levels >= 1) hp.tiles_done = &c->counters->hiz_tiles_b;
			CK(launch_hiz(hp, c->num_sms, s, &launches));
		}
		mark(E_HIZ_B);
	}
	if (timed) cudaEventRecord(c->stage_ev[E_COUNT], s); // end of frame (the per-stage events exist only with VKV_FRAME_STAGES)
	if (launches_out) *launches_out = (uint32_t)launches;
	if (out) {
		memset(out, 0, sizeof(*out));
		CK(cudaMemcpyAsync(c->h_counters, c->counters, sizeof(FrameCounters), cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		rc = check_merge_error(c); // also reported by the next vkv_sync / vkv_read_* when the caller passed no stats
		if (rc) return rc;
		out->draws = N;
		out->visible_a = c->h_counters->visible[0];
		out->occluded_a = c->h_counters->occluded[0];
		out->visible_b = c->h_counters->visible[1];
		out->tested_b = two ? c->h_counters->occluded[0] : 0;
		out->kernel_launches = (uint32_t)launches;
		out->strip_tiles_pulled = c->h_counters->strip_tiles_pulled;
		out->strip_texels_sent = c->h_counters->strip_texels_sent;
		out->hiz_tiles_b = c->h_counters->hiz_tiles_b;
		out->drain_items_a = c->h_counters->drain_seen[0]; out->drain_items_b = c->h_counters->drain_seen[1];
		observe_drain(c, c->h_counters, two);
		if (timed) {
			auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, c->stage_ev[a], c->stage_ev[b]); return ms; };
			out->total_ms = el(E_BEGIN, E_COUNT);
		}
		if (stages) {
			auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, c->stage_ev[a], c->stage_ev[b]); return ms; };
			out->clear_ms = el(E_BEGIN, E_CLEAR); out->cull_a_ms = el(E_CLEAR, E_CULL_A); out->raster_a_ms = el(E_CULL_A, E_RASTER_A);
			out->merge_a_ms = el(E_RASTER_A, E_MERGE_A); out->hiz_a_ms = el(E_MERGE_A, E_HIZ_A);
			if (two) {
				out->cull_b_ms = el(E_HIZ_A, E_CULL_B); out->raster_b_ms = el(E_CULL_B, E_RASTER_B);
				out->merge_b_ms = el(E_RASTER_B, E_MERGE_B); out->hiz_b_ms = el(E_MERGE_B, E_HIZ_B);
			}
		}
	}
	return VKV_OK;
}
} // namespace

int vkv_frame(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, uint32_t flags, vkv_stats* out) { return frame_impl(c, pc, flags, out, nullptr); }

/* Frames in flight.  The reference never waits for the frame it has just recorded: it keeps frameOverlap frames in flight, each with its
 * own camera / draw buffers, and waits on the fence of the frame it is about to reuse (application.cpp:133,642-660).  vkv_frame_submit
 * enqueues a frame plus the device->host copy of its counters into a pinned ring slot and returns; vkv_frame_wait blocks on that slot. */
int vkv_frame_submit(vkv_ctx* c, const vkv_VisbufferPushConstants* pc, uint32_t flags, uint32_t* ticket) {
	if (!c || !ticket) return c ? fail(c, VKV_ERR_INVALID, "vkv_frame_submit: NULL argument") : VKV_ERR_INVALID;
	if (flags & (VKV_FRAME_TIMED | VKV_FRAME_STAGES)) return fail(c, VKV_ERR_INVALID, "vkv_frame_submit: timed frames are blocking frames (vkv_frame)");
	vkv_ctx::Flight& f = c->flight[c->next_ticket % vkv_ctx::kFlights];
	if (f.ticket) return fail(c, VKV_ERR_LIMIT, "vkv_frame_submit: %u frames in flight already; vkv_frame_wait for ticket %u first", vkv_ctx::kFlights, f.ticket);
	uint32_t launches = 0;
	int rc = frame_impl(c, pc, flags, nullptr, &launches);
	if (rc) return rc;
	CK(cudaMemcpyAsync(f.h, c->counters, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaEventRecord(f.done, c->stream));
	f.ticket = c->next_ticket; f.draws = pc->meshletDrawCount; f.launches = launches;
	f.two = (flags & VKV_FRAME_TWO_PASS) && !(flags & VKV_FRAME_NO_CULL);
	*ticket = f.ticket;
	if (++c->next_ticket == 0) c->next_ticket = 1; // 0 marks a free slot
	return VKV_OK;
}

int vkv_frame_wait(vkv_ctx* c, uint32_t ticket, vkv_stats* out) {
	if (!c || !ticket) return VKV_ERR_INVALID;
	vkv_ctx::Flight& f = c->flight[ticket % vkv_ctx::kFlights];
	if (f.ticket != ticket) return fail(c, VKV_ERR_INVALID, "vkv_frame_wait: ticket %u is not in flight", ticket);
	CK(cudaSetDevice(c->device));
	CK(cudaEventSynchronize(f.done));
	if (out) {
		memset(out, 0, sizeof(*out));
		out->draws = f.draws;
		out->visible_a = f.h->visible[0];
		out->occluded_a = f.h->occluded[0];
		out->visible_b = f.h->visible[1];
		out->tested_b = f.two ? f.h->occluded[0] : 0;
		out->kernel_launches = f.launches;
		out->strip_tiles_pulled = f.h->strip_tiles_pulled;
		out->strip_texels_sent = f.h->strip_texels_sent;
		out->hiz_tiles_b = f.h->hiz_tiles_b;
		out->drain_items_a = f.h->drain_seen[0]; out->drain_items_b = f.h->drain_seen[1];
	}
	observe_drain(c, f.h, f.two);
	f.ticket = 0;
	return VKV_OK;
}

/* ---- draw list on the device -------------------------------------------------------------------------- */

int vkv_build_draws(vkv_ctx* c, const vkv_DrawSegment* host_segments, uint32_t n_segments, uint64_t primitiveBuffer, uint64_t* draw_buffer, uint32_t* draw_count) {
	if (!c || !draw_buffer || !draw_count || (!host_segments && n_segments)) return c ? fail(c, VKV_ERR_INVALID, "vkv_build_draws: NULL argument") : VKV_ERR_INVALID;
	if (n_segments && !primitiveBuffer) return fail(c, VKV_ERR_INVALID, "vkv_build_draws: primitiveBuffer is NULL");
	*draw_buffer = 0; *draw_count = 0;
	CK(cudaSetDevice(c->device));
	cudaStream_t s = c->stream;
	vkv_DrawSegment* dseg = nullptr;
	uint32_t* doff = nullptr;
	auto cleanup = [&]() { if (dseg) cudaFree(dseg); if (doff) cudaFree(doff); };
	cudaError_t e = cudaMalloc(&dseg, (size_t)(n_segments ? n_segments : 1) * sizeof(vkv_DrawSegment));
	if (e == cudaSuccess) e = cudaMalloc(&doff, (size_t)(n_segments + 2) * 4);
	if (e == cudaSuccess && n_segments) e = cudaMemcpyAsync(dseg, host_segments, (size_t)n_segments * sizeof(vkv_DrawSegment), cudaMemcpyHostToDevice, s);
	if (e == cudaSuccess) e = cudaMemsetAsync(doff + n_segments + 1, 0, 4, s); // overflow flag
	if (e == cudaSuccess) e = launch_segment_scan(dseg, n_segments, (const vkv_Primitive*)primitiveBuffer, doff, doff + n_segments + 1, s);
	uint32_t tail[2] = {0, 0}; // total, overflow
	if (e == cudaSuccess) e = cudaMemcpyAsync(tail, doff + n_segments, 8, cudaMemcpyDeviceToHost, s);
	if (e == cudaSuccess) e = cudaStreamSynchronize(s); // host_segments is borrowed; the total sizes the allocation
	if (e != cudaSuccess) { cleanup(); return fail(c, e == cudaErrorMemoryAllocation ? VKV_ERR_OOM : VKV_ERR_CUDA, "vkv_build_draws: %s", cudaGetErrorString(e)); }
	if (tail[1] || tail[0] > VKV_MAX_MESHLET_DRAWS) { cleanup(); return fail(c, VKV_ERR_LIMIT, "draw list exceeds 2^25 MeshletDraws (visbuffer.h.glsl:15-17)"); }
	const uint32_t total = tail[0];
	void* d = nullptr;
	const size_t bytes = (size_t)total * sizeof(vkv_MeshletDraw);
	const size_t padded = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
	e = cudaMalloc(&d, padded);
	if (e == cudaSuccess) e = launch_expand_segments(dseg, n_segments, doff, (vkv_MeshletDraw*)d, total, c->num_sms, s);
	if (e == cudaSuccess) e = cudaStreamSynchronize(s);
	cleanup();
	if (e != cudaSuccess) { if (d) cudaFree(d); return fail(c, e == cudaErrorMemoryAllocation ? VKV_ERR_OOM : VKV_ERR_CUDA, "vkv_build_draws: %s", cudaGetErrorString(e)); }
	{
		std::lock_guard<std::mutex> lock(c->mtx);
		c->allocs[(uint64_t)(uintptr_t)d] = bytes;
	}
	*draw_buffer = (uint64_t)(uintptr_t)d;
	*draw_count = total;
	return VKV_OK;
}

int vkv_download(vkv_ctx* c, uint64_t dev_addr, void* host, size_t bytes) {
	if (!c || !host) return VKV_ERR_INVALID;
	{
		std::lock_guard<std::mutex> lock(c->mtx);
		auto it = c->allocs.upper_bound(dev_addr);
		if (it == c->allocs.begin()) return fail(c, VKV_ERR_INVALID, "vkv_download: unknown address");
		--it;
		if (dev_addr + bytes > it->first + it->second) return fail(c, VKV_ERR_INVALID, "vkv_download: range exceeds the allocation");
	}
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(host, (const void*)(uintptr_t)dev_addr, bytes, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return VKV_OK;
}

/* ---- resolve ------------------------------------------------------------------------------------- */

int vkv_resolve(vkv_ctx* c, const vkv_VisbufferPushConstants* pc) {
	if (!c) return VKV_ERR_INVALID;
	if (!pc) return fail(c, VKV_ERR_INVALID, "push constants are NULL");
	CK(cudaSetDevice(c->device));
	const size_t n = (size_t)c->W * c->H;
	if (!c->color) {
		CK(cudaMalloc(&c->color, n * 4));
		CK(cudaMemsetAsync(c->color, 0, n * 4, c->stream)); // image contents before the first resolve
	}
	if (pc->meshletDrawCount == 0) return VKV_OK; // application.cpp:930: no draw buffer -> no dispatch
	if (!pc->drawBuffer || !pc->primitiveBuffer || !pc->materialBuffer) return fail(c, VKV_ERR_INVALID, "push constants hold a NULL buffer address");
	// material count = extent of the allocation the material buffer lives in (the push constants do not carry it)
	size_t bytes = 0;
	{
		std::lock_guard<std::mutex> lock(c->mtx);
		auto it = c->allocs.upper_bound(pc->materialBuffer);
		if (it != c->allocs.begin()) {
			--it;
			if (pc->materialBuffer < it->first + it->second) bytes = it->first + it->second - pc->materialBuffer;
		}
	}
	if (bytes < sizeof(vkv_Material)) return fail(c, VKV_ERR_INVALID, "materialBuffer does not point into a vkv_upload allocation");
	const uint32_t nm = (uint32_t)(bytes / sizeof(vkv_Material));
	if (nm > c->mat_cap) {
		CK(cudaStreamSynchronize(c->stream));
		uint32_t* fresh = nullptr;
		CK(cudaMalloc(&fresh, (size_t)(nm + 64) * 4));
		if (c->mat_colors) cudaFree(c->mat_colors);
		c->mat_colors = fresh; c->mat_cap = nm + 64;
	}
	CK(launch_material_colors((const vkv_Material*)pc->materialBuffer, nm, c->mat_colors, c->num_sms, c->stream));
	ResolveParams r{};
	r.vis = c->vis; r.draws = (const vkv_MeshletDraw*)pc->drawBuffer; r.primitives = (const vkv_Primitive*)pc->primitiveBuffer;
	r.matColors = c->mat_colors; r.color = c->color; r.W = c->W; r.H = c->H;
	CK(launch_resolve(r, c->num_sms, c->stream));
	return VKV_OK;
}

int vkv_read_color(vkv_ctx* c, uint32_t* host) {
	if (!c || !host) return VKV_ERR_INVALID;
	if (!c->color) return fail(c, VKV_ERR_INVALID, "vkv_read_color: no resolve has run since the targets were created");
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(host, c->color, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return VKV_OK;
}

/* ---- motion vectors (visbuffer.frag.glsl:38) ---------------------------------------------------------------------- */

int vkv_motion_vectors(vkv_ctx* c, const vkv_VisbufferPushConstants* pc) {
	if (!c) return VKV_ERR_INVALID;
	if (!pc) return fail(c, VKV_ERR_INVALID, "push constants are NULL");
	CK(cudaSetDevice(c->device));
	const size_t n = (size_t)c->W * c->H;
	if (!c->motion) CK(cudaMalloc(&c->motion, n * 4));
	if (pc->meshletDrawCount == 0) { // nothing can have been drawn: the cleared attachment (application.cpp:786-799)
		CK(cudaMemsetAsync(c->motion, 0, n * 4, c->stream));
		return VKV_OK;
	}
	if (!pc->drawBuffer || !pc->primitiveBuffer || !pc->transformBuffer || !pc->cameraBuffer) return fail(c, VKV_ERR_INVALID, "push constants hold a NULL buffer address");
	MotionParams m{};
	m.vis = c->vis; m.draws = (const vkv_MeshletDraw*)pc->drawBuffer; m.primitives = (const vkv_Primitive*)pc->primitiveBuffer;
	m.transforms = (const float*)pc->transformBuffer; m.camera = (const vkv_Camera*)pc->cameraBuffer; m.out = c->motion; m.W = c->W; m.H = c->H;
	CK(launch_motion(m, c->num_sms, c->stream));
	return VKV_OK;
}

int vkv_read_motion(vkv_ctx* c, uint16_t* host) {
	if (!c || !host) return VKV_ERR_INVALID;
	if (!c->motion) return fail(c, VKV_ERR_INVALID, "vkv_read_motion: vkv_motion_vectors has not run since the targets were created");
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(host, c->motion, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return VKV_OK;
}

/* ---- multi-GPU ------------------------------------------------------------------------------------------------ */

int vkv_set_shard(vkv_ctx* c, uint32_t first_draw, uint32_t draw_count, int enable) {
	if (!c) return VKV_ERR_INVALID;
	c->sharded = enable != 0;
	c->shard_first = enable ? first_draw : 0;
	c->shard_count = enable ? draw_count : 0;
	c->shard_block_log2 = 0; c->shard_rank = 0; c->shard_nranks = 1;
	return VKV_OK;
}

int vkv_set_shard_interleaved(vkv_ctx* c, int rank, int nranks, uint32_t block_log2) {
	if (!c) return VKV_ERR_INVALID;
	if (nranks < 1 || rank < 0 || rank >= nranks || block_log2 < 5 || block_log2 > 24)
		return fail(c, VKV_ERR_INVALID, "interleaved shard: rank %d of %d, block 2^%u (need 2^5..2^24 draws)", rank, nranks, block_log2);
	c->sharded = nranks > 1;
	c->shard_first = 0; c->shard_count = 0;
	c->shard_block_log2 = nranks > 1 ? block_log2 : 0; c->shard_rank = (uint32_t)rank; c->shard_nranks = (uint32_t)nranks;
	return VKV_OK;
}

int vkv_ipc_export(vkv_ctx* c, void* handle128) {
	if (!c || !handle128) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle layout");
	cudaIpcMemHandle_t h[2];
	CK(cudaIpcGetMemHandle(&h[0], c->vis));
	CK(cudaIpcGetMemHandle(&h[1], c->xchg));
	memcpy(handle128, h, 128);
	return VKV_OK;
}

int vkv_ipc_attach(vkv_ctx* c, int rank, int nranks, const void* handles) {
	if (!c || !handles) return VKV_ERR_INVALID;
	if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return fail(c, VKV_ERR_INVALID, "rank %d of %d out of range (max %d ranks)", rank, nranks, kMaxRanks);
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	detach_peers(c);
	c->mp.rank = rank; c->mp.nranks = nranks;
	c->mp.error = c->sync_flags + kMaxRanks;
	c->attached = true; // from here on detach_peers() closes whatever was opened
	const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)handles;
	for (int r = 0; r < nranks; ++r) {
		if (r == rank) { c->mp.vis[r] = c->vis; c->mp.flags[r] = c->sync_flags; c->mp.dirty[r] = c->dirty; c->mp.pyr[r] = c->pyramid; continue; }
		void *pv = nullptr, *pf = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&pv, h[2 * r], cudaIpcMemLazyEnablePeerAccess);
		if (e == cudaSuccess) { c->mp.vis[r] = (unsigned long long*)pv; e = cudaIpcOpenMemHandle(&pf, h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess); }
		if (e != cudaSuccess) {
			detach_peers(c);
			return fail(c, VKV_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
		}
		c->mp.flags[r] = (uint32_t*)pf;
		// same resolution on every rank (the caller's contract), hence the same layout inside the peer's exchange block
		c->mp.dirty[r] = (uint8_t*)pf + 256;
		c->mp.pyr[r] = (float*)((unsigned char*)pf + c->pyramid_offset);
	}
	// A new session starts from a clean barrier state: epoch 0 AND zeroed local slots + error word.  (Resetting only the epoch would
	// leave the previous session's epochs in the slots, and `flag - epoch >= 0` would let every barrier pass at once.)  This is
	// safe because no peer can still be signalling into these slots: a peer's stream is idle before it exports (vkv_resize /
	// vkv_ipc_detach synchronize it) and the caller exchanges the handles between that and this call.  The caller must also
	// rendezvous on the host between the LAST rank's attach and the first merged frame (multigpu.attach_peers: dist.barrier()).
	CK(cudaMemset(c->sync_flags, 0, (kMaxRanks + 1) * 4));
	c->epoch = 0;
	c->merge_err_pending = false;
	return VKV_OK;
}

int vkv_ipc_detach(vkv_ctx* c) {
	if (!c) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	detach_peers(c);
	return VKV_OK;
}

int vkv_merge(vkv_ctx* c) {
	if (!c) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	return enqueue_merge(c, nullptr);
}

int vkv_strip_owner(vkv_ctx* c, uint32_t pixel_row, int nranks) {
	if (!c) return VKV_ERR_INVALID;
	if (nranks < 1 || pixel_row >= c->H) return fail(c, VKV_ERR_INVALID, "vkv_strip_owner: row %u of %u, %d ranks", pixel_row, c->H, nranks);
	return strip_owner_of_tile_row(pixel_row / 16u, nranks);
}

int vkv_gather_strips(vkv_ctx* c) {
	if (!c) return VKV_ERR_INVALID;
	if (!c->attached) return fail(c, VKV_ERR_INVALID, "vkv_gather_strips: no peers are attached (vkv_ipc_attach)");
	CK(cudaSetDevice(c->device));
	const unsigned long long timeout_ns = 5ull * 1000 * 1000 * 1000;
	StripParams sp{};
	sp.mp = c->mp; sp.mp.n = (size_t)c->W * c->H;
	sp.W = c->W; sp.H = c->H; sp.pyr = c->pyr; sp.exact_levels = c->exact_levels;
	sp.tilesX = c->tiles_x; sp.tilesY = c->tiles_y; sp.dirtyStride = c->dirty_stride;
	CK(launch_xgpu_barrier(c->mp, ++c->epoch, timeout_ns, c->stream));   // every owner's strip is final
	CK(launch_strip_gather(sp, c->num_sms, c->stream));
	CK(launch_xgpu_barrier(c->mp, ++c->epoch, timeout_ns, c->stream));   // nobody overwrites a strip a peer is still reading
	c->merge_used = true;
	return VKV_OK;
}

int vkv_hash(vkv_ctx* c, int what, int rank, int nranks, uint64_t* out) {
	if (!c || !out) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	unsigned long long* d = (unsigned long long*)c->tmp_count; // 256-byte scratch
	CK(cudaMemsetAsync(d, 0, 8, c->stream));
	if (what == 0) {
		if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, VKV_ERR_INVALID, "vkv_hash: rank %d of %d", rank, nranks);
		CK(launch_hash_owned(c->vis, c->W, c->H, rank, nranks, d, c->num_sms, c->stream));
	} else if (what == 1) { // the pyramid as u64 words (its allocation is padded with a zero float to an even count)
		CK(launch_hash64((const unsigned long long*)c->pyramid, 0, ((size_t)c->pyr.total + 1) / 2, d, c->num_sms, c->stream));
	} else return fail(c, VKV_ERR_INVALID, "vkv_hash: what must be 0 (visbuffer rows of a rank) or 1 (pyramid)");
	unsigned long long h = 0;
	CK(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	*out = h;
	return check_merge_error(c);
}

int vkv_read_visbuffer64(vkv_ctx* c, uint64_t* host) {
	if (!c || !host) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(host, c->vis, (size_t)c->W * c->H * 8, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return check_merge_error(c);
}

static int read_split(vkv_ctx* c, uint32_t* ids, float* depth) {
	const size_t n = (size_t)c->W * c->H;
	CK(cudaSetDevice(c->device));
	if (!c->tmp_ids) CK(cudaMalloc(&c->tmp_ids, n * 4));
	if (!c->tmp_depth) CK(cudaMalloc(&c->tmp_depth, n * 4));
	CK(launch_split_vis(c->vis, n, ids ? c->tmp_ids : nullptr, depth ? c->tmp_depth : nullptr, c->num_sms, c->stream));
	if (ids) CK(cudaMemcpyAsync(ids, c->tmp_ids, n * 4, cudaMemcpyDeviceToHost, c->stream));
	if (depth) CK(cudaMemcpyAsync(depth, c->tmp_depth, n * 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return check_merge_error(c);
}
int vkv_read_ids(vkv_ctx* c, uint32_t* host) { return (c && host) ? read_split(c, host, nullptr) : VKV_ERR_INVALID; }
int vkv_read_depth(vkv_ctx* c, float* host) { return (c && host) ? read_split(c, nullptr, host) : VKV_ERR_INVALID; }

uint32_t vkv_pyramid_floats(vkv_ctx* c) { return c ? c->pyr.total : 0; }

int vkv_read_hiz_mip(vkv_ctx* c, uint32_t mip, float* host, uint32_t* w, uint32_t* h) {
	if (!c) return VKV_ERR_INVALID;
	if (mip >= c->pyr.levels) return fail(c, VKV_ERR_INVALID, "mip %u >= %u levels", mip, c->pyr.levels);
	if (w) *w = c->pyr.w[mip];
	if (h) *h = c->pyr.h[mip];
	if (host) {
		CK(cudaSetDevice(c->device));
		CK(cudaMemcpyAsync(host, c->pyramid + c->pyr.off[mip], (size_t)c->pyr.w[mip] * c->pyr.h[mip] * 4, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return VKV_OK;
}

int vkv_read_pyramid(vkv_ctx* c, float* host, uint32_t floats) {
	if (!c || !host) return VKV_ERR_INVALID;
	if (floats != c->pyr.total) return fail(c, VKV_ERR_INVALID, "pyramid holds %u floats, caller passed %u", c->pyr.total, floats);
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(host, c->pyramid, (size_t)floats * 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return check_merge_error(c);
}

int vkv_write_pyramid(vkv_ctx* c, const float* host, uint32_t floats) {
	if (!c || !host) return VKV_ERR_INVALID;
	if (floats != c->pyr.total) return fail(c, VKV_ERR_INVALID, "pyramid holds %u floats, caller passed %u", c->pyr.total, floats);
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(c->pyramid, host, (size_t)floats * 4, cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return VKV_OK;
}

int vkv_read_visible(vkv_ctx* c, int pass, uint32_t* draw_ids, uint32_t cap, uint32_t* n) {
	if (!c || pass < 0 || pass > 1 || !n) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(c->h_counters, c->counters, sizeof(FrameCounters), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	const uint32_t cnt = c->h_counters->visible[pass];
	*n = cnt;
	if (draw_ids && cnt) {
		if (cap < cnt) return fail(c, VKV_ERR_INVALID, "vkv_read_visible: capacity %u < %u survivors", cap, cnt);
		CK(cudaMemcpyAsync(draw_ids, c->list_visible[pass], (size_t)cnt * 4, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return VKV_OK;
}

int vkv_read_status(vkv_ctx* c, int pass, uint8_t* status, uint32_t n) {
	if (!c || pass < 0 || pass > 1 || !status) return VKV_ERR_INVALID;
	if (!c->status_valid[pass]) return fail(c, VKV_ERR_INVALID, "status of pass %d was not recorded (pass VKV_FRAME_STATUS)", pass);
	if (n > c->cap_draws) return fail(c, VKV_ERR_INVALID, "n exceeds the draw capacity");
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(status, c->status[pass], n, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return VKV_OK;
}

int vkv_event_record(vkv_ctx* c, int slot) {
	if (!c || slot < 0 || slot >= 16) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	CK(cudaEventRecord(c->events[slot], c->stream));
	return VKV_OK;
}

int vkv_event_elapsed(vkv_ctx* c, int from, int to, float* ms) {
	if (!c || !ms || from < 0 || from >= 16 || to < 0 || to >= 16) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	CK(cudaEventSynchronize(c->events[to]));
	CK(cudaEventElapsedTime(ms, c->events[from], c->events[to]));
	return VKV_OK;
}

int vkv_flush_l2(vkv_ctx* c, size_t bytes) {
	if (!c) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	if (bytes > c->flush_bytes) {
		CK(cudaStreamSynchronize(c->stream));
		if (c->flush_buf) cudaFree(c->flush_buf);
		c->flush_buf = nullptr; c->flush_bytes = 0; // scratch only: nothing else refers to it, so free-then-allocate is safe
		CK(cudaMalloc(&c->flush_buf, bytes));
		c->flush_bytes = bytes;
	}
	CK(launch_fill32((uint32_t*)c->flush_buf, bytes / 4, 0u, c->num_sms, c->stream));
	return VKV_OK;
}

uint64_t vkv_visbuffer64_ptr(vkv_ctx* c) { return c ? (uint64_t)(uintptr_t)c->vis : 0; }
uint64_t vkv_motion_ptr(vkv_ctx* c) { return c ? (uint64_t)(uintptr_t)c->motion : 0; }

int vkv_alloc(vkv_ctx* c, size_t bytes, uint64_t* dev_addr) {
	if (!c || !dev_addr) return c ? fail(c, VKV_ERR_INVALID, "vkv_alloc: NULL argument") : VKV_ERR_INVALID;
	std::lock_guard<std::mutex> lock(c->mtx);
	CK(cudaSetDevice(c->device));
	void* d = nullptr;
	const size_t padded = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
	CK(cudaMalloc(&d, padded));
	CK(cudaMemsetAsync(d, 0, padded, c->stream));
	c->allocs[(uint64_t)(uintptr_t)d] = bytes;
	*dev_addr = (uint64_t)(uintptr_t)d;
	return VKV_OK;
}

// ---- EXT_meshopt_compression decode (kernels in meshopt.cu) ---------------------------------------------
struct vkv_meshopt_plan {
	MeshoptPlan p{};
	uint32_t n_views = 0;
	unsigned long long src_need = 0, dst_need = 0;   // extents the views reach into the two buffers
	std::vector<void*> dev;                          // every device allocation of the plan
};

int vkv_meshopt_plan_create(vkv_ctx* c, const vkv_MeshoptView* views, uint32_t n, vkv_meshopt_plan** out) {
	if (!c || !out || (!views && n)) return c ? fail(c, VKV_ERR_INVALID, "vkv_meshopt_plan_create: NULL argument") : VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	std::vector<MeshoptStream> vs, is, fs;
	std::vector<unsigned long long> elemFirst;
	std::vector<uint32_t> blockStream;
	unsigned long long planes = 0, felems = 0, srcNeed = 0, dstNeed = 0;
	for (uint32_t i = 0; i < n; ++i) {
		const vkv_MeshoptView& v = views[i];
		MeshoptStream st{};
		st.src_off = v.src_offset; st.src_size = v.src_size; st.dst_off = v.dst_offset;
		st.count = v.count; st.stride = v.stride; st.mode = v.mode; st.filter = v.filter; st.view = i;
		if (v.dst_offset % 4) return fail(c, VKV_ERR_INVALID, "meshopt view %u: dst_offset must be a multiple of 4", i);
		if (v.src_size >= (1ull << 30)) return fail(c, VKV_ERR_LIMIT, "meshopt view %u: compressed views of 1 GiB and more are not supported", i);
		if (v.mode == VKV_MESHOPT_ATTRIBUTES) {
			if (v.stride == 0 || v.stride > 256 || v.stride % 4) return fail(c, VKV_ERR_INVALID, "meshopt view %u: byteStride %u (must be a multiple of 4, <= 256)", i, v.stride);
			if ((v.filter == VKV_MESHOPT_FILTER_OCT && v.stride != 4 && v.stride != 8) || (v.filter == VKV_MESHOPT_FILTER_QUAT && v.stride != 8) || v.filter > 3)
				return fail(c, VKV_ERR_INVALID, "meshopt view %u: filter %u does not fit byteStride %u", i, v.filter, v.stride);
			uint32_t bs = (8192u / v.stride) & ~15u; // vertexcodec.cpp:116-126
			st.block_size = bs < 256u ? bs : 256u;
			const unsigned long long nb = ((unsigned long long)v.count + st.block_size - 1) / st.block_size;
			if (blockStream.size() + nb >= (1ull << 31) || planes + nb * v.stride >= (1ull << 32)) return fail(c, VKV_ERR_LIMIT, "meshopt plan: too many vertex blocks");
			st.first_block = (uint32_t)blockStream.size();
			st.plane_base = (uint32_t)planes;
			planes += nb * v.stride;
			blockStream.insert(blockStream.end(), (size_t)nb, (uint32_t)vs.size());
			vs.push_back(st);
			if (v.filter) {
				elemFirst.push_back(felems);
				felems += v.filter == VKV_MESHOPT_FILTER_EXP ? (unsigned long long)v.count * (v.stride / 4) : v.count;
				fs.push_back(st);
			}
		} else if (v.mode == VKV_MESHOPT_TRIANGLES || v.mode == VKV_MESHOPT_INDICES) {
			if (v.stride != 2 && v.stride != 4) return fail(c, VKV_ERR_INVALID, "meshopt view %u: index size %u (must be 2 or 4)", i, v.stride);
			if (v.mode == VKV_MESHOPT_TRIANGLES && v.count % 3) return fail(c, VKV_ERR_INVALID, "meshopt view %u: triangle view with count %u", i, v.count);
			if (v.filter) return fail(c, VKV_ERR_INVALID, "meshopt view %u: filters apply to attribute views only", i);
			is.push_back(st);
		} else return fail(c, VKV_ERR_INVALID, "meshopt view %u: mode %u", i, v.mode);
		srcNeed = std::max(srcNeed, (unsigned long long)v.src_offset + v.src_size);
		dstNeed = std::max(dstNeed, (unsigned long long)v.dst_offset + (unsigned long long)v.count * v.stride);
	}
	vkv_meshopt_plan* pl = new vkv_meshopt_plan();
	pl->n_views = n; pl->src_need = srcNeed; pl->dst_need = dstNeed;
	auto put = [&](const void* host, size_t bytes, const void** devOut) -> cudaError_t {
		void* d = nullptr;
		cudaError_t e = cudaMalloc(&d, bytes ? bytes : 4);
		if (e != cudaSuccess) return e;
		pl->dev.push_back(d);
		if (host && bytes) e = cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice);
		*devOut = d;
		return e;
	};
	cudaError_t e = cudaSuccess;
	const void* d = nullptr;
	if (e == cudaSuccess) { e = put(vs.data(), vs.size() * sizeof(MeshoptStream), &d); pl->p.vertexStreams = (const MeshoptStream*)d; pl->p.nVertexStreams = (uint32_t)vs.size(); }
	if (e == cudaSuccess) { e = put(is.data(), is.size() * sizeof(MeshoptStream), &d); pl->p.indexStreams = (const MeshoptStream*)d; pl->p.nIndexStreams = (uint32_t)is.size(); }
	if (e == cudaSuccess) { e = put(fs.data(), fs.size() * sizeof(MeshoptStream), &d); pl->p.filtered = (const MeshoptStream*)d; pl->p.nFiltered = (uint32_t)fs.size(); }
	if (e == cudaSuccess) { e = put(elemFirst.data(), elemFirst.size() * 8, &d); pl->p.elemFirst = (const unsigned long long*)d; pl->p.filterElems = felems; }
	if (e == cudaSuccess) { e = put(blockStream.data(), blockStream.size() * 4, &d); pl->p.blockStream = (const uint32_t*)d; pl->p.nBlocks = (uint32_t)blockStream.size(); }
	if (e == cudaSuccess) { e = put(nullptr, (size_t)planes * 4, &d); pl->p.planeOff = (uint32_t*)d; }
	if (e == cudaSuccess) { e = put(nullptr, (size_t)planes, &d); pl->p.totals = (uint32_t*)d; }  // planes / 4 words
	if (e == cudaSuccess) { e = put(nullptr, (size_t)n * 4, &d); pl->p.status = (int32_t*)d; }
	if (e != cudaSuccess) {
		vkv_meshopt_plan_destroy(c, pl);
		return fail(c, e == cudaErrorMemoryAllocation ? VKV_ERR_OOM : VKV_ERR_CUDA, "vkv_meshopt_plan_create: %s", cudaGetErrorString(e));
	}
	*out = pl;
	return VKV_OK;
}

int vkv_meshopt_run(vkv_ctx* c, vkv_meshopt_plan* pl, uint64_t src_dev, size_t src_bytes, uint64_t dst_dev, size_t dst_bytes) {
	if (!c || !pl) return c ? fail(c, VKV_ERR_INVALID, "vkv_meshopt_run: NULL argument") : VKV_ERR_INVALID;
	if (pl->src_need > src_bytes) return fail(c, VKV_ERR_INVALID, "vkv_meshopt_run: a view reaches byte %llu of a %zu-byte source", pl->src_need, src_bytes);
	if (pl->dst_need > dst_bytes) return fail(c, VKV_ERR_INVALID, "vkv_meshopt_run: a view reaches byte %llu of a %zu-byte destination", pl->dst_need, dst_bytes);
	if (pl->n_views == 0) return VKV_OK;
	if ((!src_dev && pl->src_need) || (!dst_dev && pl->dst_need) || dst_dev % 4) return fail(c, VKV_ERR_INVALID, "vkv_meshopt_run: bad buffer address");
	CK(cudaSetDevice(c->device));
	CK(launch_meshopt_decode(pl->p, (const uint8_t*)(uintptr_t)src_dev, (uint8_t*)(uintptr_t)dst_dev, c->num_sms, c->stream, nullptr));
	return VKV_OK;
}

int vkv_meshopt_results(vkv_ctx* c, vkv_meshopt_plan* pl, int32_t* results) {
	if (!c || !pl || (!results && pl->n_views)) return c ? fail(c, VKV_ERR_INVALID, "vkv_meshopt_results: NULL argument") : VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	if (pl->n_views) CK(cudaMemcpyAsync(results, pl->p.status, (size_t)pl->n_views * 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return VKV_OK;
}

void vkv_meshopt_plan_destroy(vkv_ctx* c, vkv_meshopt_plan* pl) {
	if (!pl) return;
	if (c) cudaSetDevice(c->device);
	for (void* d : pl->dev) cudaFree(d);
	delete pl;
}

// ---- accessor conversions (kernels in accessors.cu) ---------------------------------------------------------------------
int vkv_assemble_vertices(vkv_ctx* c, uint64_t positions_dev, uint32_t type, int normalized, uint32_t byte_stride, uint32_t count, uint64_t* vertices_dev) {
	if (!c || !vertices_dev) return c ? fail(c, VKV_ERR_INVALID, "vkv_assemble_vertices: NULL argument") : VKV_ERR_INVALID;
	if (type != 5120 && type != 5121 && type != 5122 && type != 5123 && type != 5126) return fail(c, VKV_ERR_INVALID, "vkv_assemble_vertices: componentType %u", type);
	if (count && !positions_dev) return fail(c, VKV_ERR_INVALID, "vkv_assemble_vertices: NULL source");
	const uint32_t cs = (type == 5120 || type == 5121) ? 1u : (type == 5126 ? 4u : 2u);
	if (byte_stride == 0) byte_stride = 3 * cs;
	if (byte_stride < 3 * cs) return fail(c, VKV_ERR_INVALID, "vkv_assemble_vertices: byteStride %u < element size", byte_stride);
	int rc = vkv_alloc(c, (size_t)count * 24, vertices_dev);
	if (rc) return rc;
	CK(cudaSetDevice(c->device));
	CK(launch_assemble_vertices((const uint8_t*)(uintptr_t)positions_dev, (int)type, normalized, byte_stride, count, (void*)(uintptr_t)*vertices_dev, c->num_sms, c->stream));
	return VKV_OK;
}

int vkv_widen_indices(vkv_ctx* c, uint64_t indices_dev, uint32_t type, uint32_t count, uint64_t* indices32_dev) {
	if (!c || !indices32_dev) return c ? fail(c, VKV_ERR_INVALID, "vkv_widen_indices: NULL argument") : VKV_ERR_INVALID;
	if (type != 5121 && type != 5123 && type != 5125) return fail(c, VKV_ERR_INVALID, "vkv_widen_indices: componentType %u", type);
	if (count && !indices_dev) return fail(c, VKV_ERR_INVALID, "vkv_widen_indices: NULL source");
	int rc = vkv_alloc(c, (size_t)count * 4, indices32_dev);
	if (rc) return rc;
	CK(cudaSetDevice(c->device));
	CK(launch_widen_indices((const uint8_t*)(uintptr_t)indices_dev, (int)type, count, (uint32_t*)(uintptr_t)*indices32_dev, c->num_sms, c->stream));
	return VKV_OK;
}

// ---- meshlet partition + bounds (kernels in meshlets.cu) -------------------------------------------------
int vkv_build_meshlets(vkv_ctx* c, const vkv_MeshletBuildInput* in, uint32_t n, uint32_t vertex_stride, uint32_t max_vertices, uint32_t max_triangles,
                       vkv_MeshletBuildOutput* out) {
	if (!c || !out || (!in && n)) return c ? fail(c, VKV_ERR_INVALID, "vkv_build_meshlets: NULL argument") : VKV_ERR_INVALID;
	if (max_vertices < 3 || max_vertices > 64 || max_triangles < 1 || max_triangles > 252) return fail(c, VKV_ERR_INVALID, "vkv_build_meshlets: limits %u / %u (max 64 / 252)", max_vertices, max_triangles);
	if (vertex_stride < 12 || vertex_stride % 4) return fail(c, VKV_ERR_INVALID, "vkv_build_meshlets: vertex_stride %u", vertex_stride);
	CK(cudaSetDevice(c->device));
	std::vector<MeshletBuildPrim> prims(n);
	std::vector<uint32_t> triFirst(n + 1, 0);
	std::vector<MeshletBuildSeg> segs;
	unsigned long long total = 0;
	for (uint32_t i = 0; i < n; ++i) {
		if (in[i].index_count % 3) return fail(c, VKV_ERR_INVALID, "vkv_build_meshlets: primitive %u has %u indices", i, in[i].index_count);
		if (in[i].index_count && (!in[i].indices || !in[i].vertices)) return fail(c, VKV_ERR_INVALID, "vkv_build_meshlets: primitive %u: NULL buffer", i);
		prims[i] = MeshletBuildPrim{(const uint32_t*)(uintptr_t)in[i].indices, (const uint8_t*)(uintptr_t)in[i].vertices, in[i].vertex_count, 0};
		const uint32_t T = in[i].index_count / 3;
		triFirst[i] = (uint32_t)total;
		for (uint32_t t = 0; t < T; t += kMeshletSeg)
			segs.push_back(MeshletBuildSeg{(uint32_t)total + t, (uint32_t)total + std::min(T, t + kMeshletSeg), i, t == 0 ? 1u : 0u});
		total += T;
		if (total >= (1ull << 30)) return fail(c, VKV_ERR_LIMIT, "vkv_build_meshlets: more than 2^30 triangles in one call");
	}
	triFirst[n] = (uint32_t)total;
	for (uint32_t i = 0; i < n; ++i) out[i] = vkv_MeshletBuildOutput{0, 0, 0, 0, 0, 0, 0};
	if (total == 0) return VKV_OK;

	std::vector<void*> scratch;
	auto release = [&]() { for (void* d : scratch) cudaFree(d); };
	auto dev = [&](size_t bytes, const void* host, void** outp) -> cudaError_t {
		void* d = nullptr;
		cudaError_t e = cudaMalloc(&d, bytes ? bytes : 4);
		if (e != cudaSuccess) return e;
		scratch.push_back(d);
		if (host) e = cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, c->stream);
		*outp = d;
		return e;
	};
#define MBCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { release(); return fail(c, e_ == cudaErrorMemoryAllocation ? VKV_ERR_OOM : VKV_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
	MeshletBuildJob j{};
	j.nPrims = n; j.totalTris = (uint32_t)total; j.nSegs = (uint32_t)segs.size(); j.maxV = max_vertices; j.maxT = max_triangles; j.vertexStride = vertex_stride;
	void* d = nullptr;
	MBCK(dev(prims.size() * sizeof(MeshletBuildPrim), prims.data(), &d)); j.prims = (const MeshletBuildPrim*)d;
	MBCK(dev(triFirst.size() * 4, triFirst.data(), &d)); j.triFirst = (const uint32_t*)d;
	MBCK(dev(segs.size() * sizeof(MeshletBuildSeg), segs.data(), &d)); j.segs = (const MeshletBuildSeg*)d;
	MBCK(dev((size_t)total, nullptr, &d)); j.len = (uint8_t*)d;
	MBCK(dev((size_t)total, nullptr, &d)); j.ucnt = (uint8_t*)d;
	MBCK(dev(segs.size() * max_triangles * sizeof(MeshletBuildSegEntry), nullptr, &d)); j.table = (MeshletBuildSegEntry*)d;
	MBCK(dev(segs.size() * sizeof(MeshletBuildSegState), nullptr, &d)); j.state = (MeshletBuildSegState*)d;
	MBCK(dev((size_t)(n + 1) * 12, nullptr, &d)); j.primBase = (uint32_t*)d;
	MBCK(dev(4, nullptr, &d)); j.badIndex = (uint32_t*)d;
	MBCK(cudaMemsetAsync(j.badIndex, 0, 4, c->stream));
	MBCK(launch_meshlet_scan(j, c->stream));
	std::vector<uint32_t> base((size_t)(n + 1) * 3);
	uint32_t bad = 0;
	MBCK(cudaMemcpyAsync(base.data(), j.primBase, base.size() * 4, cudaMemcpyDeviceToHost, c->stream));
	MBCK(cudaMemcpyAsync(&bad, j.badIndex, 4, cudaMemcpyDeviceToHost, c->stream));
	MBCK(cudaStreamSynchronize(c->stream));
	if (bad) { // nothing has dereferenced a vertex yet (the scan reads indices only): reject before the emit pass would
		release();
		return fail(c, VKV_ERR_INVALID, "vkv_build_meshlets: primitive %u holds an index >= its vertex_count %u", bad - 1, in[bad - 1].vertex_count);
	}
	const uint32_t M = base[n * 3], V = base[n * 3 + 1], B = base[n * 3 + 2];
	MBCK(dev((size_t)M * sizeof(MeshletBuildRecord), nullptr, &d)); j.rec = (MeshletBuildRecord*)d;
	// outputs: three allocations shared by the primitives of this call, registered like vkv_upload's
	void* outs[3] = {nullptr, nullptr, nullptr};
	const size_t sizes[3] = {(size_t)M * sizeof(vkv_Meshlet), (size_t)V * 4, (size_t)B};
	for (int k = 0; k < 3; ++k) {
		const size_t padded = ((sizes[k] ? sizes[k] : 1) + 255) & ~(size_t)255;
		cudaError_t e = cudaMalloc(&outs[k], padded);
		if (e == cudaSuccess) e = cudaMemsetAsync(outs[k], 0, padded, c->stream);
		if (e != cudaSuccess) {
			for (int q = 0; q <= k; ++q) if (outs[q]) cudaFree(outs[q]);
			release();
			return fail(c, e == cudaErrorMemoryAllocation ? VKV_ERR_OOM : VKV_ERR_CUDA, "vkv_build_meshlets: output allocation: %s", cudaGetErrorString(e));
		}
	}
	cudaError_t e = launch_meshlet_emit(j, M, (vkv_Meshlet*)outs[0], (uint32_t*)outs[1], (uint8_t*)outs[2], c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	release();
	if (e != cudaSuccess) {
		for (int k = 0; k < 3; ++k) cudaFree(outs[k]);
		return fail(c, VKV_ERR_CUDA, "vkv_build_meshlets: %s", cudaGetErrorString(e));
	}
#undef MBCK
	{
		std::lock_guard<std::mutex> lock(c->mtx);
		for (int k = 0; k < 3; ++k) c->allocs[(uint64_t)(uintptr_t)outs[k]] = sizes[k];
	}
	for (uint32_t i = 0; i < n; ++i) {
		const uint32_t* b0 = &base[i * 3];
		const uint32_t* b1 = &base[(i + 1) * 3];
		out[i].meshlets = (uint64_t)(uintptr_t)outs[0] + (uint64_t)b0[0] * sizeof(vkv_Meshlet);
		out[i].vertex_indices = (uint64_t)(uintptr_t)outs[1] + (uint64_t)b0[1] * 4;
		out[i].triangles = (uint64_t)(uintptr_t)outs[2] + b0[2];
		out[i].meshlet_count = b1[0] - b0[0]; out[i].vertex_index_count = b1[1] - b0[1]; out[i].triangle_bytes = b1[2] - b0[2];
	}
	return VKV_OK;
}

int vkv_selftest_division(vkv_ctx* c, uint64_t seed, uint32_t iters_per_thread, uint64_t* tested, uint64_t* mismatches) {
	if (!c || !tested || !mismatches) return VKV_ERR_INVALID;
	CK(cudaSetDevice(c->device));
	unsigned long long* d = (unsigned long long*)c->tmp_count; // 256-byte scratch
	CK(cudaMemsetAsync(d, 0, 16, c->stream));
	CK(launch_division_selftest(seed, iters_per_thread, d, kNegZero2, c->num_sms, c->stream));
	unsigned long long h[2] = {0, 0};
	CK(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	*mismatches = h[0]; *tested = h[1];
	return VKV_OK;
}

} // extern "C"
