// kernels.cuh — parameter blocks and launchers of the sm_100a kernels (one .cu per stage).
#pragma once
#include "common.cuh"

struct CullParams {
	const vkv_MeshletDraw* draws;
	const float* transforms;
	const vkv_Primitive* primitives;
	const vkv_Camera* camera;
	const float* pyramid;
	PyramidDesc pyr;
	uint32_t n;                  // number of draws this launch tests when in_list is NULL (upper bound of the work size otherwise)
	uint32_t first;              // first MeshletDraw of this GPU's shard (in_list == NULL): draw ids stay global
	uint32_t shard_block_log2;   // interleaved sharding: blocks of 2^k draws dealt round-robin to the ranks (0 = contiguous)
	uint32_t shard_rank, shard_nranks;
	uint32_t total;              // meshletDrawCount (bound for interleaved ids)
	const uint32_t* in_list;     // pass B: the draws pass A rejected by occlusion; NULL = all draws [0,n)
	const uint32_t* in_count;    // device count of in_list
	uint32_t* out_visible;
	uint32_t* out_occluded;      // may be NULL
	FrameCounters* counters;
	uint8_t* status;             // may be NULL
	int pass;                    // 0 = A, 1 = B
	int vp_select;               // 0 = prevOcclusionViewProjection (reference), 1 = viewProjection (pass B)
	int skip_hiz;                // frustum only
	unsigned long long neg_zero2; // the fp32 pair (-0.0, -0.0) = 0x8000000080000000: see mul2 in cull.cu (must arrive at run time)
	ulonglong2* clear_ptr;       // fused visbuffer clear (vkv_frame pass A): the launch's blocks also store clear_value over
	size_t clear_n2;             // clear_n2 16-byte words — HBM write traffic issued under the compute-bound cull
	unsigned long long clear_value;
	float* xf_mvp;               // fused per-transform prologue (vkv_frame pass A): mvp / determinant sign of xf_n transforms for the raster
	uint32_t* xf_det;
	uint32_t xf_n;
	int skip_frustum;            // pass B inside vkv_frame: every input draw already passed this frame's frustum test in pass A
	const unsigned long long* cone_table; // optional cone cull: per primitive the address of its vkv_MeshletCone[] (NULL = off)
	const float4* xf_eye;        // per transform: camera position in mesh space (transform_prologue), valid BEFORE this launch
	uint32_t* reset_ptr;         // vkv_frame: the raster's queue counters (FrameCounters::big_next ..), zeroed by this launch instead of a memset
	uint32_t reset_words;        // node of their own — so that the raster launch can follow this one directly (programmatic dependent launch)
	uint4* zero_ptr;             // strip mode: the dirty-tile flags of both passes, zeroed by the same launch
	uint32_t zero_n16;
};

// A set-up triangle (24.8 fixed-point vertices, positive area) — what the rasteriser's inner loops consume.
struct Tri {
	int ax, ay, bx, by, cx, cy;      // snapped vertices, 24.8 fixed point, area2 > 0
	int xmin, xmax, ymin, ymax;      // pixel bbox clipped to the viewport
	long long area2;
	float za, dzb, dzc, invA;
	uint32_t id;
	uint32_t small;                  // vertex extent <= 2^14 sub-pixels in x and y: every edge value fits in int32
};
// Large-triangle queue entry: the triangle plus its slice of the global tile-work index space
struct BigTri {
	Tri t;
	uint32_t tileBase;               // first tile-work index of this triangle
	uint32_t tilesX, tilesY;         // bbox extent in kBigTileW x kBigTileH tiles
	uint32_t pad;
};

// A triangle some vertex of which is outside the near / far / guard-band planes: clip coordinates, waiting for the clipper
struct ClipTri {
	float4 a, b, c;                  // clip x, y, z, w
	uint32_t id;
	uint32_t pad[3];
};
constexpr int kBigSlotShift = 40;    // FrameCounters::big_cursor = records << 40 | tile-work items (raster.cu push_big)
struct RasterParams {
	const vkv_MeshletDraw* draws;
	const float* transforms;
	const vkv_Primitive* primitives;
	const vkv_Material* materials;
	const vkv_Camera* camera;
	const uint32_t* list;        // MeshletDraw indices to rasterise
	const uint32_t* count;       // device count
	uint32_t* work;              // work-stealing cursor (zeroed per launch)
	unsigned long long* vis;     // W*H 64-bit keys
	uint32_t W, H;
	const float* mvp;            // per transform: viewProjection * transform (launch_prepare_transforms)
	const uint32_t* detNeg;      // per transform: determinant(transform) < 0
	BigTri* big;                 // large-triangle queue (filled by raster_kernel, drained by raster_big_kernel)
	uint32_t bigCap;
	unsigned long long* bigCursor;
	uint32_t* bigNext;
	ClipTri* clip;               // clip queue (filled by raster_kernel, drained by raster_big_kernel's first phase)
	uint32_t clipCap;
	uint32_t* clipCount;
	uint32_t* clipNext;
	uint32_t* overflow;          // set when either queue was full
	uint32_t* drainBarrier;
	uint32_t* slowWork;
	uint32_t* drainSeen;         // statistics (NULL = none): the drain kernel stores how much it found queued
	unsigned long long neg_zero2; // the fp32 pair (-0.0, -0.0), see common.cuh mul2 (must arrive at run time)
	const vkv_QuantizedPositions* qtable; // optional: per primitive, its POSITION accessor in 16-bit form (NULL = read the f32 Vertex records)
	uint8_t* dirty;              // strip mode: one byte per 64x16-pixel tile, set for every tile a drawn triangle's bbox touches (NULL otherwise)
	uint32_t dirtyTilesX;
	uint32_t markLimit;          // tiles are marked only while *count <= markLimit (strip mode: ~0u; pass-B pyramid rebuild: kPartialHizLimit)
};

struct HizParams {
	const unsigned long long* vis;
	float* pyramid;
	PyramidDesc pyr;
	uint32_t W, H;
	uint32_t exact_levels;       // leading mips whose source is exactly 2x (handled by the tiled kernel), <= 4
	uint32_t* done;              // FrameCounters::hiz_done (zero between launches)
	int split_tail;              // diagnosis only: run the small mips as a second launch
	const uint8_t* tile_dirty;   // pass-B rebuild: one byte per 64x16-pixel tile, set by the pass-B rasteriser for every tile it may have drawn
	                             // into; clean tiles keep the exact mips the pass-A build stored a moment ago (NULL = rebuild every tile)
	const uint32_t* dirty_count; // ... valid only if *dirty_count (pass B's survivor count) <= dirty_limit: the rasteriser stops marking above it
	uint32_t dirty_limit;
	uint32_t* tiles_done;        // statistics (NULL = none): += the number of tiles this launch reduced
	// strip mode: the small-mip tail first waits for every rank's "my strip's mips are stored everywhere" signal (xgpu.cuh)
	const uint32_t* wait_flags;  // this rank's flag slots (NULL = no wait)
	uint32_t wait_epoch;
	int wait_ranks;
	uint32_t* wait_error;
	unsigned long long wait_timeout_ns;
};

cudaError_t launch_cull(const CullParams& p, int num_sms, cudaStream_t stream, bool after_hiz = false); // after_hiz: programmatic dependent launch
cudaError_t launch_iota(const CullParams& p, uint32_t* out, uint32_t* count, int num_sms, cudaStream_t stream);
cudaError_t launch_prepare_transforms(const float* transforms, const vkv_Camera* camera, uint32_t n, float* mvp, uint32_t* detNeg, float4* eye,
                                      int num_sms, cudaStream_t stream);
cudaError_t launch_raster(const RasterParams& p, int num_sms, cudaStream_t stream, bool after_cull = false, bool small_drain = false); // raster_kernel + raster_big_kernel; after_cull: programmatic dependent launch behind cull_kernel
cudaError_t launch_hiz(const HizParams& p, int num_sms, cudaStream_t stream, int* launches);
cudaError_t launch_fill64(unsigned long long* dst, size_t n, unsigned long long value, int num_sms, cudaStream_t stream);
cudaError_t launch_fill32(uint32_t* dst, size_t n, uint32_t value, int num_sms, cudaStream_t stream);
cudaError_t launch_split_vis(const unsigned long long* vis, size_t n, uint32_t* ids, float* depth, int num_sms, cudaStream_t stream);

// ---- multi-GPU visbuffer min-merge over NVLink peer memory (merge.cu) ------------------------------------------------
constexpr int kMaxRanks = 16;
struct MergeParams {
	unsigned long long* vis[kMaxRanks];   // every rank's W*H visbuffer (index == rank; [rank] is the local one)
	uint32_t* flags[kMaxRanks];           // every rank's barrier slots (kMaxRanks u32 each)
	float* pyr[kMaxRanks];                // every rank's pyramid (strip mode: strip owners store their mips into all of them)
	uint8_t* dirty[kMaxRanks];            // every rank's dirty-tile flags: [pass][tile], one byte per 64x16-pixel tile
	int rank, nranks;
	size_t n;                             // W*H
	uint32_t* error;                      // local: set to 1 when a barrier timed out
};
cudaError_t launch_xgpu_barrier(const MergeParams& p, uint32_t epoch, unsigned long long timeout_ns, cudaStream_t stream);
cudaError_t launch_merge_min(const MergeParams& p, int num_sms, cudaStream_t stream);

// ---- strip mode (strips.cu): rank r owns every n-th row of 64x16-pixel tiles ---------------------------------------------------
struct StripParams {
	MergeParams mp;
	uint32_t W, H;
	PyramidDesc pyr;
	uint32_t exact_levels;
	uint32_t tilesX, tilesY;              // 64x16-pixel tiles
	uint32_t dirtyStride;                 // bytes per pass in the dirty array (tilesX * tilesY rounded up to 16)
	int pass;                             // which dirty set the peers' contributions are read from
	uint32_t* stats;                      // [0] += tiles pulled from peers, [1] += pyramid texels stored to peers (may be NULL)
	// the two cross-GPU barriers of an exchange ride inside the kernels instead of being launches of their own:
	uint32_t epoch_in;                    // every block waits until all ranks have signalled this epoch (their raster pass is complete)
	uint32_t epoch_out;                   // signalled by the block that finishes last (this rank's strip is merged, its mips stored everywhere)
	uint32_t* done;                       // last-block ticket (zero between launches)
	unsigned long long timeout_ns;
};
// ownership: tile row tr (16 pixel rows) belongs to rank tr % n — strips INTERLEAVED over the screen, so that every rank gets its
// share of the busy rows and of the empty ones without any coordination (a contiguous split left the rank holding the sky idle and
// the one holding the densest rows pulling three times the average)
__host__ __device__ inline int strip_owner_of_tile_row(uint32_t tileRow, int nranks) { return (int)(tileRow % (uint32_t)nranks); }
__host__ __device__ inline uint32_t strip_tile_rows_owned(uint32_t tilesY, int rank, int nranks) {
	return (uint32_t)rank < tilesY ? (tilesY - (uint32_t)rank + (uint32_t)nranks - 1u) / (uint32_t)nranks : 0u;
}
cudaError_t launch_strip_merge_hiz(const StripParams& p, int num_sms, cudaStream_t stream);
cudaError_t launch_strip_gather(const StripParams& p, int num_sms, cudaStream_t stream);
cudaError_t launch_hash64(const unsigned long long* data, size_t first, size_t count, unsigned long long* out, int num_sms, cudaStream_t stream);
cudaError_t launch_hash_owned(const unsigned long long* vis, uint32_t W, uint32_t H, int rank, int nranks, unsigned long long* out, int num_sms, cudaStream_t stream);
cudaError_t launch_hiz_tail(const HizParams& p, cudaStream_t stream);

// ---- visbuffer resolve (resolve.cu) ---------------------------------------------------------------------------------
struct ResolveParams {
	const unsigned long long* vis;
	const vkv_MeshletDraw* draws;
	const vkv_Primitive* primitives;
	const uint32_t* matColors;   // per material: RGBA8 of fromLinear(albedoFactor)
	uint32_t* color;             // W*H RGBA8 (R in the low byte)
	uint32_t W, H;
};
cudaError_t launch_material_colors(const vkv_Material* materials, uint32_t n, uint32_t* out, int num_sms, cudaStream_t stream);
cudaError_t launch_resolve(const ResolveParams& p, int num_sms, cudaStream_t stream);

// ---- motion vectors from the finished visbuffer (motion.cu) --------------------------------------------------------
struct MotionParams {
	const unsigned long long* vis;
	const vkv_MeshletDraw* draws;
	const vkv_Primitive* primitives;
	const float* transforms;     // mat4[], column-major
	const vkv_Camera* camera;    // viewProjection / prevViewProjection
	uint32_t* out;               // W*H texels of R16G16_SFLOAT (x in the low half)
	uint32_t W, H;
};
cudaError_t launch_motion(const MotionParams& p, int num_sms, cudaStream_t stream);

// ---- device-side draw-list generation (drawlist.cu) ---------------------------------------------------------------
cudaError_t launch_segment_scan(const vkv_DrawSegment* seg, uint32_t n, const vkv_Primitive* prims, uint32_t* offsets, uint32_t* overflow, cudaStream_t stream);
cudaError_t launch_expand_segments(const vkv_DrawSegment* seg, uint32_t n, const uint32_t* offsets, vkv_MeshletDraw* draws, uint32_t capacity,
                                   int num_sms, cudaStream_t stream);

// ---- EXT_meshopt_compression decode (meshopt.cu) -------------------------------------------------------------------
struct MeshoptStream {               // one compressed buffer view, as the kernels see it
	unsigned long long src_off, src_size, dst_off;
	uint32_t count, stride;          // elements and bytes per element (vertex size, or index size 2 / 4)
	uint32_t mode, filter;           // vkv_MeshoptView::mode / ::filter
	uint32_t view;                   // index into the caller's view array (status slot)
	uint32_t block_size;             // vertex streams: vertices per block (vertexcodec.cpp:116-126)
	uint32_t first_block;            // vertex streams: global index of the stream's first block
	uint32_t plane_base;             // vertex streams: first entry of the stream in planeOff (a multiple of 4; totals use plane_base / 4)
};
struct MeshoptPlan {                 // device-resident description of a decode job (built once per asset)
	const MeshoptStream* vertexStreams; uint32_t nVertexStreams;
	const MeshoptStream* indexStreams; uint32_t nIndexStreams;
	const MeshoptStream* filtered; uint32_t nFiltered;      // views with a filter, with the element prefix below
	const unsigned long long* elemFirst; unsigned long long filterElems;
	const uint32_t* blockStream; uint32_t nBlocks;          // vertex block -> vertex stream
	uint32_t* planeOff;              // scratch: start of every byte plane of every block (u32 per plane)
	uint32_t* totals;                // scratch: per block and 4 planes, the block's delta total, then its carry
	int32_t* status;                 // per view: meshoptimizer's return code
};
cudaError_t launch_meshopt_decode(const MeshoptPlan& p, const uint8_t* src, uint8_t* dst, int num_sms, cudaStream_t stream, int* launches);

// ---- meshlet partition + bounds (meshlets.cu) ----------------------------------------------------------------------
constexpr uint32_t kMeshletSeg = 2048;   // triangles per chain segment
struct MeshletBuildPrim { const uint32_t* indices; const uint8_t* vertices; uint32_t vertexCount; uint32_t pad; };   // vertexCount 0 = unknown (indices not range-checked)
struct MeshletBuildSeg { uint32_t start, end; uint32_t prim; uint32_t first_of_prim; };   // global triangle range [start, end) inside one primitive
struct MeshletBuildSegEntry { uint32_t meshlets, vertices, bytes, exit; };                 // per (segment, entry offset)
struct MeshletBuildSegState { uint32_t entry, meshlets, vertices, bytes; };                // where the chain enters a segment, output bases there
struct MeshletBuildRecord { uint32_t start, vertices, bytes; };                            // per meshlet: first triangle (global), output offsets (global)
struct MeshletBuildJob {
	const MeshletBuildPrim* prims; uint32_t nPrims;
	const uint32_t* triFirst;        // [nPrims + 1] global index of each primitive's first triangle
	uint32_t totalTris;
	const MeshletBuildSeg* segs; uint32_t nSegs;
	uint32_t maxV, maxT, vertexStride;
	uint8_t* len; uint8_t* ucnt;     // per triangle: size and distinct-vertex count of the greedy meshlet starting there
	uint32_t* badIndex;              // set to primitive + 1 when an index >= that primitive's vertexCount is found (meshopt_buildMeshlets asserts this)
	MeshletBuildSegEntry* table;     // [nSegs * maxT]
	MeshletBuildSegState* state;     // [nSegs]
	uint32_t* primBase;              // [(nPrims + 1) * 3] meshlets / vertex indices / triangle bytes before each primitive (last = totals)
	MeshletBuildRecord* rec;         // [total meshlets]
};
cudaError_t launch_meshlet_scan(const MeshletBuildJob& j, cudaStream_t stream);   // len + segment + chain -> primBase
cudaError_t launch_meshlet_emit(const MeshletBuildJob& j, uint32_t nMeshlets, vkv_Meshlet* meshlets, uint32_t* meshletVertices, uint8_t* meshletTriangles,
                                cudaStream_t stream);

// ---- accessor conversions (accessors.cu) ----------------------------------------------------------------------------
cudaError_t launch_assemble_vertices(const uint8_t* src, int type, int normalized, uint32_t stride, uint32_t count, void* vertices, int num_sms, cudaStream_t stream);
cudaError_t launch_widen_indices(const uint8_t* src, int type, uint32_t count, uint32_t* out, int num_sms, cudaStream_t stream);

// ---- arithmetic self checks (selftest.cu) -------------------------------------------------------------------------
cudaError_t launch_division_selftest(uint64_t seed, uint32_t iters, unsigned long long* counters2, unsigned long long negZero2, int num_sms, cudaStream_t stream);
