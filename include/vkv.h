/* vkv.h — C ABI of the B200 geometry hot path (libvkv.so): meshlet cull -> visibility-buffer raster -> HiZ pyramid.
 *
 * The reference has no plugin/FFI interface; the seam this library replaces is the code region in
 * Application::run() that records the mesh-shader draw and the hiz_reduce dispatch loop, plus the target creation
 * behind it (paths relative to the upstream tree):
 *
 *   vkv_create / vkv_resize   initVisbufferPass (application.cpp:181-270) + initHiZReductionPass (:472-529),
 *                             updateRenderResolution (:578-602)
 *   vkv_upload / vkv_update   createMeshBuffers + staging copies (assets.cpp:233-286,385-424), World::addAsset
 *                             (world.cpp:89-178), rebuildDrawBuffer / updateTransformBuffer (world.cpp:267-290,321-344),
 *                             Camera mapped write (camera.cpp:109,180-193)
 *   vkv_frame                 "Visbuffer pass" (application.cpp:763-867: clears :782,:807, push constants :849-859,
 *                             vkCmdDrawMeshTasksEXT :861) + "HiZ reduction" (:951-1003)
 *   vkv_cull                  shaders/visbuffer/visbuffer.task.glsl:25-76 (+ culling.h.glsl:8-56)
 *   vkv_raster                shaders/visbuffer/visbuffer.mesh.glsl:30-104 + fixed-function raster/depth state
 *                             (application.cpp:326-340) + visbuffer.frag.glsl:36
 *   vkv_hiz                   shaders/hiz_reduce.comp.glsl:21-31 + dispatch loop application.cpp:964-1000
 *   vkv_read_*                what the resolve pass / next frame's task shader read (application.cpp:917-949)
 *
 * Errors: the reference throws vulkan_error from vk::checkResult (include/vulkan/vk.hpp:27-61); here every call returns
 * 0 or a negative vkv_status and never throws; vkv_last_error() gives the message.
 * Threading: one context per GPU; frame/stage calls are single-threaded per context (the reference records frames on
 * one thread); vkv_upload/vkv_free are internally locked (the reference uploads from worker threads).
 * All `host` pointers are borrowed for the duration of the call. No torch / CUDA types appear in any signature.
 */
#ifndef VKV_H
#define VKV_H

#include "vkv_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vkv_ctx vkv_ctx;

enum vkv_status {
	VKV_OK = 0,
	VKV_ERR_CUDA = -1,        /* a CUDA runtime call failed (message has the cudaError string) */
	VKV_ERR_INVALID = -2,     /* bad argument */
	VKV_ERR_NO_DEVICE = -3,   /* no CUDA device / driver: there is NO CPU fallback */
	VKV_ERR_OOM = -4,
	VKV_ERR_LIMIT = -5        /* meshletDrawCount > 2^25 (visbuffer.h.glsl:15-17) */
};

/* vkv_frame flags */
enum {
	VKV_FRAME_ONE_PASS = 0,        /* reference behaviour: cull against the PREVIOUS pyramid with prevOcclusionViewProjection */
	VKV_FRAME_TWO_PASS = 1 << 0,   /* extension: + re-test pass-A occlusion rejects against the fresh pyramid */
	VKV_FRAME_NO_HIZ = 1 << 1,     /* camera->freezeCullingMatrix: skip the pyramid rebuild (application.cpp:951) */
	VKV_FRAME_STATUS = 1 << 2,     /* also write the per-draw status bytes (parity / debugging) */
	VKV_FRAME_TIMED = 1 << 3,      /* fill total_ms of vkv_stats (CUDA events around the frame; syncs the stream at frame end) */
	VKV_FRAME_NO_CULL = 1 << 4,    /* rasterise every MeshletDraw (debug; what the task shader does with culling disabled) */
	VKV_FRAME_MERGE = 1 << 5,      /* multi-GPU: min-merge the visbuffer with the attached peers before each pyramid build */
	VKV_FRAME_MERGE_STRIPS = 1 << 7, /* multi-GPU: screen-strip ownership — each rank pulls the tiles its peers drew into the tile rows it owns,
	                                  builds their exact mips and stores the changed texels into every rank's pyramid; the merged
	                                  visbuffer stays distributed (vkv_strip_owner), the pyramid is complete everywhere */
	VKV_FRAME_CONE_CULL = 1 << 8,  /* extension, OFF in parity mode: reject meshlets whose normal cone faces away from the camera before the
	                                  occlusion test (needs vkv_set_cone_table).  Removes only meshlets none of whose triangles the mesh
	                                  shader's facing test would keep: the visbuffer is unchanged, the visible lists get shorter */
	VKV_FRAME_STAGES = 1 << 6      /* with VKV_FRAME_TIMED: also the per-stage *_ms fields (an event between two launches; it keeps the
	                                  pass-B cull from overlapping the pyramid's tail, so total_ms is a few us higher than without) */
};

/* per-draw status byte (VKV_FRAME_STATUS) — same values as the oracle's */
enum { VKV_ST_FRUSTUM_CULLED = 0, VKV_ST_OCCLUDED = 1, VKV_ST_VISIBLE = 2, VKV_ST_NOT_TESTED = 3,
       VKV_ST_CONE_CULLED = 0x80 /* VKV_FRAME_CONE_CULL: rejected by the normal cone (class bits 0: culled before the occlusion test) */ };

typedef struct vkv_stats {
	uint32_t draws;              /* meshletDrawCount */
	uint32_t visible_a, occluded_a, visible_b, tested_b;
	float clear_ms, cull_a_ms, raster_a_ms, hiz_a_ms, cull_b_ms, raster_b_ms, hiz_b_ms, total_ms; /* total_ms: VKV_FRAME_TIMED; the rest: + VKV_FRAME_STAGES */
	uint32_t kernel_launches;    /* kernels of this library launched by the call */
	float merge_a_ms, merge_b_ms; /* VKV_FRAME_TIMED + VKV_FRAME_STAGES + VKV_FRAME_MERGE / VKV_FRAME_MERGE_STRIPS (strip mode: barrier + pull-merge +
	                                 exact mips + all-gather + barrier; hiz_*_ms is then the small-mip tail alone) */
	uint32_t strip_tiles_pulled; /* VKV_FRAME_MERGE_STRIPS, both passes: (64x16-pixel tile, peer) pairs this rank pulled over NVLink, 8 KB each */
	uint32_t strip_texels_sent;  /* ... and pyramid texels it stored into its peers, 4 B each */
	uint32_t hiz_tiles_b;        /* two-pass frames: 64x16-pixel tiles the SECOND pyramid build reduced — all of them, or, after a small pass B,
	                                only the tiles that pass drew into (the first build always reduces every tile) */
	uint32_t drain_items_a, drain_items_b; /* what the drain kernel behind each pass's rasteriser found queued (clip triangles + large-triangle
	                                records + 1 if a queue overflowed).  The library sizes the next frames' drain launches from it: after two
	                                observed frames with an empty pass the launch shrinks (an empty drain is pure launch latency) */
} vkv_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int vkv_create(vkv_ctx** out, int cuda_device, uint32_t width, uint32_t height);
int vkv_resize(vkv_ctx*, uint32_t width, uint32_t height);
void vkv_destroy(vkv_ctx*);
const char* vkv_last_error(vkv_ctx*);  /* ctx may be NULL: message of the last failed vkv_create on this thread */
/* run on an existing CUDA stream (cudaStream_t passed as void*); NULL restores the context's own stream */
int vkv_set_stream(vkv_ctx*, void* cuda_stream);
int vkv_sync(vkv_ctx*);

/* ---- device memory ---------------------------------------------------------------------------------------- */
int vkv_upload(vkv_ctx*, const void* host, size_t bytes, uint64_t* dev_addr);       /* alloc + copy; address is 256-byte aligned */
int vkv_update(vkv_ctx*, uint64_t dev_addr, const void* host, size_t bytes);        /* rewrite (camera, transforms) — async on the ctx stream */
int vkv_free(vkv_ctx*, uint64_t dev_addr);
int vkv_download(vkv_ctx*, uint64_t dev_addr, void* host, size_t bytes);            /* read back part of a vkv_upload / vkv_build_draws buffer */

/* ---- per frame -------------------------------------------------------------------------------------------- */
int vkv_frame(vkv_ctx*, const vkv_VisbufferPushConstants* pc, uint32_t flags, vkv_stats* out);
/* Frames in flight.  The reference never waits for the frame it has just recorded: Application::run() keeps frameOverlap frames in
 * flight, each with its own camera buffer / draw buffer / command pool, and waits on the fence of the frame slot it is about to reuse
 * (application.cpp:133,153,167,642; camera.cpp:86, world.cpp:5-6).  The same shape here:
 *   vkv_update_staged  copies from PINNED host memory on the context's upload stream, i.e. beside the kernels of the frame before;
 *                      the next vkv_frame / vkv_frame_submit (and vkv_sync) waits for it.  The individually callable stages and the
 *                      vkv_read_* / vkv_download family do NOT: pair those with vkv_update, or call vkv_sync in between.  The caller owns the
 *                      hazard the reference owns too: the destination must not be read by a frame still in flight (use one camera /
 *                      transform buffer per frame slot), and the pinned source must stay untouched until that frame has been waited for.
 *   vkv_frame_submit   vkv_frame without the wait: enqueues the frame and the device->host copy of its counters, returns a ticket (!= 0).
 *                      At most 4 tickets may be outstanding (VKV_ERR_LIMIT).  VKV_FRAME_TIMED / _STAGES are refused (blocking by nature).
 *   vkv_frame_wait     blocks until that frame and its counters have arrived; fills `out` (may be NULL) exactly as vkv_frame does
 *                      (the *_ms fields stay 0) and releases the ticket.  It does not drain the stream: a multi-GPU barrier timeout of a
 *                      submitted frame is reported by the next vkv_sync / vkv_read_* / blocking vkv_frame, as for vkv_frame(out = NULL). */
int vkv_update_staged(vkv_ctx*, uint64_t dev_addr, const void* pinned_host, size_t bytes);
int vkv_frame_submit(vkv_ctx*, const vkv_VisbufferPushConstants* pc, uint32_t flags, uint32_t* ticket);
int vkv_frame_wait(vkv_ctx*, uint32_t ticket, vkv_stats* out);
/* Side buffer of the optional cone cull (VKV_FRAME_CONE_CULL): device address of a table with one entry per primitive — the address
 * of that primitive's vkv_MeshletCone[meshletCount] (vkv_abi.h; vkvh_scene_upload_cones builds both).  0 removes it. */
int vkv_set_cone_table(vkv_ctx*, uint64_t table_dev_addr);
/* KHR_mesh_quantization positions dequantised in registers (extension): device address of a vkv_QuantizedPositions table
 * with one entry per primitive (vkv_abi.h; vkvh_scene_upload_quantized builds it).  The rasteriser then reads 8 bytes per vertex —
 * the accessor's own int16 data — instead of the expanded 24-byte Vertex record, and converts exactly as fastgltf does on the
 * host: the image is bit-identical.  Culling is unaffected (it reads meshlet bounds only).  0 removes the table. */
int vkv_set_quantized_positions(vkv_ctx*, uint64_t table_dev_addr);
/* individually callable stages (per-stage timing / parity).  pass: 0 = A (reference), 1 = B (two-pass extension) */
int vkv_clear(vkv_ctx*);
int vkv_cull(vkv_ctx*, const vkv_VisbufferPushConstants* pc, int pass, uint32_t flags, uint32_t* n_visible);
int vkv_raster(vkv_ctx*, const vkv_VisbufferPushConstants* pc, int pass);
int vkv_hiz(vkv_ctx*);
/* rasterise an explicit MeshletDraw index list (host pointer) — test hook */
int vkv_raster_list(vkv_ctx*, const vkv_VisbufferPushConstants* pc, const uint32_t* draw_ids, uint32_t n);

/* ---- draw list on the device.  Replaces World::rebuildDrawBuffer (world.cpp:230-293): the host passes one
 * vkv_DrawSegment per (mesh-node, primitive) in traversal order (host pointer, 8 B each) instead of 12 B per MeshletDraw;
 * the library expands them (length = Primitive.meshletCount read on the device) into a MeshletDraw[] it owns.  The buffer is
 * byte-identical to the reference's host-built list; *draw_buffer can go straight into the push constants, vkv_free releases it.
 * VKV_ERR_LIMIT if the list would exceed 2^25 draws. --------------------------------------------------------------------- */
int vkv_build_draws(vkv_ctx*, const vkv_DrawSegment* host_segments, uint32_t n_segments, uint64_t primitiveBuffer,
                    uint64_t* draw_buffer, uint32_t* draw_count);

/* ---- EXT_meshopt_compression decode on the device.  Replaces CompressedBufferDataAdapter::ExecuteRange
 * (assets.cpp:111-171): per compressed buffer view, meshopt_decodeVertexBuffer / meshopt_decodeIndexBuffer /
 * meshopt_decodeIndexSequence followed by meshopt_decodeFilterOct / Quat / Exp.  A view is fastgltf's CompressedBufferView
 * (mc.mode, mc.filter, mc.count, mc.byteStride, mc.byteOffset, mc.byteLength) plus where its decoded bytes go.  The plan is
 * built once per asset (descriptors + scratch on the device); vkv_meshopt_run enqueues the decode of the whole asset on the
 * context's stream: src_dev = the compressed glTF buffer (vkv_upload), dst_dev = a vkv_alloc'ed buffer whose sub-ranges can
 * go straight into Primitive's buffer addresses.  results[i] is meshoptimizer's return code for view i (0, -1 bad header or
 * version, -2 truncated, -3 trailing bytes); a failed view leaves its output range undefined and does not affect the others.
 * Filters follow meshoptimizer's scalar definitions (vertexfilter.cpp:75-160; see oracle/meshopt_decode.cpp). ------------- */
enum { VKV_MESHOPT_ATTRIBUTES = 0, VKV_MESHOPT_TRIANGLES = 1, VKV_MESHOPT_INDICES = 2 };               /* fastgltf MeshoptCompressionMode */
enum { VKV_MESHOPT_FILTER_NONE = 0, VKV_MESHOPT_FILTER_OCT = 1, VKV_MESHOPT_FILTER_QUAT = 2, VKV_MESHOPT_FILTER_EXP = 3 }; /* MeshoptCompressionFilter */
typedef struct vkv_MeshoptView {
	uint32_t mode, filter;
	uint32_t count, stride;         /* mc.count, mc.byteStride */
	uint64_t src_offset, src_size;  /* mc.byteOffset, mc.byteLength within the compressed buffer */
	uint64_t dst_offset;            /* where count * stride decoded bytes go within the destination buffer (multiple of 4) */
} vkv_MeshoptView;
typedef struct vkv_meshopt_plan vkv_meshopt_plan;
int vkv_meshopt_plan_create(vkv_ctx*, const vkv_MeshoptView* host_views, uint32_t n_views, vkv_meshopt_plan** out);
int vkv_meshopt_run(vkv_ctx*, vkv_meshopt_plan*, uint64_t src_dev, size_t src_bytes, uint64_t dst_dev, size_t dst_bytes); /* async */
int vkv_meshopt_results(vkv_ctx*, vkv_meshopt_plan*, int32_t* results);   /* n_views return codes of the last run (syncs the stream) */
void vkv_meshopt_plan_destroy(vkv_ctx*, vkv_meshopt_plan*);
int vkv_alloc(vkv_ctx*, size_t bytes, uint64_t* dev_addr);                /* zero-filled device buffer, same bookkeeping as vkv_upload */

/* ---- accessor conversions on the device.  Replace the host loops that fill glsl::Vertex and the u32 index vector
 * (assets.cpp:308-320: fastgltf::iterateAccessor<glm::vec3> on POSITION, copyFromAccessor<uint32_t> on the indices): per
 * component fastgltf's convertComponent<float, T> (tools.hpp:266-289) — float(x), or max(float(x) / float(T max), -1) for
 * normalized integers (KHR_mesh_quantization).  component_type is the glTF enum (5120 BYTE, 5121 UNSIGNED_BYTE, 5122 SHORT,
 * 5123 UNSIGNED_SHORT, 5125 UNSIGNED_INT, 5126 FLOAT); byte_stride 0 = tightly packed.  The source is a device address
 * (e.g. inside the vkv_meshopt_run destination); the result is a new allocation (vkv_free): glsl::Vertex[count] with only
 * `position` set, as the reference leaves it, or uint32[count]. -------------------------------------------------------- */
int vkv_assemble_vertices(vkv_ctx*, uint64_t positions_dev, uint32_t component_type, int normalized, uint32_t byte_stride, uint32_t count,
                          uint64_t* vertices_dev);
int vkv_widen_indices(vkv_ctx*, uint64_t indices_dev, uint32_t component_type, uint32_t count, uint64_t* indices32_dev);

/* ---- meshlet partition + bounds on the device.  Stands where PrimitiveProcessingTask::processPrimitive
 * builds meshlets and their AABBs on the host (assets.cpp:322-373).  Per primitive: a u32 triangle-list index buffer and the
 * vertex buffer (position = three floats at the start of every `vertex_stride` bytes: glsl::Vertex, 24 B), both device
 * addresses.  Output per primitive: Meshlet[] in the reference's 36-byte layout with aabbExtents / aabbCenter computed as
 * assets.cpp:349-372 does, the meshlet vertex-index list (u32) and the meshlet triangle bytes (each meshlet padded to 4) —
 * the three arrays glsl::Primitive points at.  The partition is meshoptimizer's SCAN partition (meshopt_buildMeshletsScan:
 * triangles in index order), byte-identical to that function; the reference's meshopt_buildMeshlets + meshopt_optimizeMeshlet
 * order heuristics are not reproduced (no golden output exists for them; any valid partition renders the same image).
 * max_vertices <= 64, max_triangles <= 252 (the reference uses 64 / 124).  All primitives of a call share three allocations:
 * release out[0].meshlets, out[0].vertex_indices and out[0].triangles with vkv_free.
 * Trust boundary: indices come from untrusted glTF.  vertex_count > 0 makes the call range-check every index on the device
 * and fail with VKV_ERR_INVALID before any vertex is dereferenced (meshopt_buildMeshlets asserts the same, clusterizer.cpp:45);
 * vertex_count == 0 skips the check (the caller vouches for the indices).  The per-frame path trusts primitiveIndex /
 * meshletIndex / transformIndex / materialIndex of the buffers it is handed, exactly as the reference's shaders do. ------ */
typedef struct vkv_MeshletBuildInput { uint64_t indices, vertices; uint32_t index_count, vertex_count; } vkv_MeshletBuildInput;
typedef struct vkv_MeshletBuildOutput {
	uint64_t meshlets, vertex_indices, triangles;                       /* device addresses of this primitive's arrays */
	uint32_t meshlet_count, vertex_index_count, triangle_bytes, reserved;
} vkv_MeshletBuildOutput;
int vkv_build_meshlets(vkv_ctx*, const vkv_MeshletBuildInput* host_inputs, uint32_t n, uint32_t vertex_stride, uint32_t max_vertices,
                       uint32_t max_triangles, vkv_MeshletBuildOutput* host_outputs);

/* ---- resolve: visbuffer -> RGBA8 colour image.  Replaces shaders/visbuffer/visbuffer_resolve.comp.glsl:17-41
 * and its dispatch (application.cpp:917-949); the push constants' drawBuffer / primitiveBuffer / materialBuffer are the fields
 * of the reference's VisbufferResolvePushConstants (visbuffer.h.glsl:49-56).  Texel = R | G<<8 | B<<16 | A<<24, sRGB-encoded
 * fromLinear(material.albedoFactor) (srgb.h.glsl:26-32); untouched / undrawn pixels are 0. ------------------------------- */
int vkv_resolve(vkv_ctx*, const vkv_VisbufferPushConstants* pc);
int vkv_read_color(vkv_ctx*, uint32_t* host);                      /* W*H RGBA8 */

/* ---- motion vectors: the visbuffer pass's second colour attachment (application.cpp:250-267 R16G16_SFLOAT "Motion vectors", cleared
 * to 0 at :786-799, handed to the upscaler at :1086-1118), derived from the finished visbuffer.  Replaces visbuffer.frag.glsl:38 and the
 * position / prevPosition varyings of visbuffer.mesh.glsl:44-45,61-63: per covered pixel the triangle its id names is fetched again and
 *   ((prevPosition.xy / prevPosition.w) * 0.5 - 0.5) - ((position.xy / position.w) * 0.5 - 0.5)
 * is evaluated at the pixel centre with perspective-correct interpolation (csrc/motion_core.h states the arithmetic); uncovered pixels
 * hold the clear value.  pc->cameraBuffer must hold the frame's Camera (prevViewProjection = the previous frame's viewProjection,
 * camera.cpp:181).  Call it after the frame whose visbuffer it reads; with range sharding, after vkv_gather_strips / vkv_merge.
 * Texel = half(x) | half(y) << 16. ------------------------------------------------------------------------------------------- */
int vkv_motion_vectors(vkv_ctx*, const vkv_VisbufferPushConstants* pc);
int vkv_read_motion(vkv_ctx*, uint16_t* host);                     /* W*H*2 halves */
uint64_t vkv_motion_ptr(vkv_ctx*);                                 /* device address of the W*H texels for a GPU-resident consumer; 0 before the first pass */

/* ---- multi-GPU: one process (and one context) per GPU; a single huge view is sharded by MeshletDraw range and the
 * per-GPU 64-bit visbuffers are min-merged over NVLink peer memory (BASELINE config 5).  The reference is
 * single-GPU: there is no call site to cite, only the data contract — drawIndex stays the index into the GLOBAL list
 * (visbuffer.h.glsl:15-16), so the merged image is bit-identical to a single-GPU frame. -------------------------------- */
/* this GPU culls / rasterises MeshletDraws [first_draw, first_draw + draw_count) of pc->drawBuffer; enable = 0 restores the whole list */
int vkv_set_shard(vkv_ctx*, uint32_t first_draw, uint32_t draw_count, int enable);
/* load-balanced variant: the list is cut into blocks of 2^block_log2 draws dealt round-robin, rank r owns blocks r, r+n, ...
 * (occlusion makes contiguous halves of a list very unequal in surviving work); nranks = 1 switches sharding off */
int vkv_set_shard_interleaved(vkv_ctx*, int rank, int nranks, uint32_t block_log2);
/* 128-byte opaque handle (two cudaIpcMemHandle_t: visbuffer + barrier flags) to pass to the other ranks (any transport) */
int vkv_ipc_export(vkv_ctx*, void* handle128);
/* handles = nranks * 128 bytes, in rank order (the caller's own entry is ignored); opens the peers' buffers */
int vkv_ipc_attach(vkv_ctx*, int rank, int nranks, const void* handles);
int vkv_ipc_detach(vkv_ctx*);
/* all ranks call it at the same point of their stream: barrier, fused reduce-scatter + all-gather u64 min, barrier */
int vkv_merge(vkv_ctx*);
/* Strip ownership (VKV_FRAME_MERGE_STRIPS).  The screen is cut into rows of 64x16-pixel tiles dealt round-robin: tile row t (pixel rows
 * 16t .. 16t+15) belongs to rank t % nranks — interleaved, so every rank gets its share of busy and of empty rows.  After a strip-mode
 * frame a rank's visbuffer holds the MERGED keys in the rows it owns (its other rows hold what this rank drew); the pyramid is
 * complete and identical on every rank.  vkv_strip_owner returns the owning rank of a pixel row (or a negative vkv_status).
 * vkv_gather_strips (a collective: all ranks call it at the same point) pulls the rows a rank does not own from their owners, so
 * that every rank holds the whole merged image. */
int vkv_strip_owner(vkv_ctx*, uint32_t pixel_row, int nranks);
int vkv_gather_strips(vkv_ctx*);
/* order-independent 64-bit digest computed on the device: what = 0: the visbuffer keys of the rows `rank` owns among `nranks`
 * (nranks = 1: the whole image); what = 1: the whole pyramid (rank / nranks ignored).  Equal data at equal positions give equal
 * digests on any GPU (parity checks between a sharded frame and the same frame rendered by one GPU). */
int vkv_hash(vkv_ctx*, int what, int rank, int nranks, uint64_t* out);

/* ---- results (blocking device->host copies on the ctx stream) ------------------------------------------- */
int vkv_read_visbuffer64(vkv_ctx*, uint64_t* host);                /* W*H keys: (~floatBits(depth) << 32) | packVisBuffer */
int vkv_read_ids(vkv_ctx*, uint32_t* host);                        /* W*H, == the reference's R32_UINT attachment */
int vkv_read_depth(vkv_ctx*, float* host);                         /* W*H, == the reference's D32_SFLOAT attachment */
int vkv_read_hiz_mip(vkv_ctx*, uint32_t mip, float* host, uint32_t* w, uint32_t* h);
int vkv_read_pyramid(vkv_ctx*, float* host, uint32_t floats);      /* all mips, contiguous (layout: vkv_abi.h) */
int vkv_write_pyramid(vkv_ctx*, const float* host, uint32_t floats);/* test hook: preset "previous frame" pyramid */
int vkv_read_visible(vkv_ctx*, int pass, uint32_t* draw_ids, uint32_t cap, uint32_t* n);   /* survivors, unordered */
int vkv_read_status(vkv_ctx*, int pass, uint8_t* status, uint32_t n);                      /* needs VKV_FRAME_STATUS */
uint32_t vkv_pyramid_floats(vkv_ctx*);

/* ---- measurement helpers -------------------------------------------------------------------------------------- */
/* CUDA events on the ctx stream: record slot i (0..15); elapsed ms between two recorded slots (syncs on the later one) */
int vkv_event_record(vkv_ctx*, int slot);
int vkv_event_elapsed(vkv_ctx*, int from, int to, float* ms);
/* write `bytes` of a scratch buffer (> L2) to evict the working set between timed iterations */
int vkv_flush_l2(vkv_ctx*, size_t bytes);
/* device pointer of the 64-bit visbuffer (for the multi-GPU min-merge; see vkv_merge_*) */
uint64_t vkv_visbuffer64_ptr(vkv_ctx*);
/* arithmetic self check: the kernels divide clip.xyz by clip.w with ONE refined reciprocal per vertex / AABB corner
 * (culling.h.glsl:49-51 and the perspective divide after visbuffer.mesh.glsl:61 are three `/` by the same w); this runs
 * that routine against the IEEE `/` operator on 148*8*256*iters_per_thread pseudo-random operand quadruples on the GPU and
 * returns how many quotients were compared and how many differed in any bit (must be 0). */
int vkv_selftest_division(vkv_ctx*, uint64_t seed, uint32_t iters_per_thread, uint64_t* tested, uint64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif
