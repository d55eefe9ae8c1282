/* oracle.h — CPU restatement of the reference's per-frame geometry path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker (or as the timed CPU baseline), never on the product path.
 *
 * PARITY UNPINNED only for the BITS of interpolated depth (Vulkan leaves that arithmetic to the implementation) and for the sentence
 * "a MIN-reduction fetch returns the minimum over the texels with non-zero weight" (Vulkan specification; nothing here can execute it).
 * The reference ships no tests, golden images or known-answer vectors for this path and its shaders cannot be run
 * in this image (no Vulkan loader / lavapipe / glslang).  Two things pin the rest (tests/test_oracle.py, tests/test_llvmpipe.py):
 *   (1) the fixed-function stages, for which no reference CODE exists — triangle coverage (pixel centres, top-left rule, 8 sub-pixel
 *       bits, clipping), the >= depth test, the LINEAR sampler's texel footprint, mip selection — against LLVMPIPE, the rasteriser and
 *       texture unit underneath lavapipe (Mesa 18.1.9 inside this image's Nsight Compute, driven without an X server through
 *       oracle/llvmpipe/fakex11.c): identical coverage and ids, identical footprints at every coordinate hiz_reduce samples; depth
 *       within ~1e-7.  llvmpipe's outputs are committed as tests/golden/llvmpipe.npz;
 *   (1b) the reference's shader text executed AS GLSL by that Mesa's compiler (tests/test_llvmpipe_glsl.py): mesh-shader clip positions,
 *       both determinants and gl_CullPrimitiveEXT BIT-IDENTICAL to orc_mesh_shader on BASELINE cfg 1-4, task-shader classes identical on
 *       every MeshletDraw — the arithmetic policy below is exactly Mesa's lowering of the reference's GLSL;
 *   (2) the reference's own text, compiled as C++ against its glm by oracle/build_ref.sh into oracle/_ref/:
 *   - the whole per-draw decision of the task shader except the texture fetch: culling.h.glsl (isAabbInFrustum,
 *     getWorldSpaceAabbExtent, aabbPositions, projectAabb), visbuffer.task.glsl:50-52,56-61,64 — on every MeshletDraw of all
 *     five BASELINE configs at full size the oracle's class differs from the glm-evaluated reference only on draws it flags
 *     ORC_AMBIG_* / ORC_CROSSES_CAMERA (0-7 draws in 10 M, all flagged);
 *   - the mesh shader's arithmetic: visbuffer.mesh.glsl:44 (mvp), :61 (gl_Position), :71 (transformDet), :86-98 (the facing decision
 *     on determinant(mat3(v0.xyw, v1.xyw, v2.xyw))) — clip positions agree with the glm-evaluated lines to < 4e-7 of the vertex's largest
 *     coordinate, determinants within 2.2 % of orc_mesh_shader's first-order noise bound, gl_CullPrimitiveEXT differs only by sign
 *     flips inside that bound (none unflagged).  NOTE what that test measures about the REFERENCE: the determinant of three nearly
 *     parallel (x, y, w) vectors cancels catastrophically for small distant triangles — 0.9 % (cfg 3) to 3.9 % (cfg 5) of the facing
 *     decisions depend on how the compiler associates the products.  The oracle's association (below) is the definition here;
 *   - the HiZ reduce pass except the texture fetch: hiz_reduce.comp.glsl:21-31 (main) run for every invocation of the dispatches
 *     application.cpp:964-979 records (mip count :472-473, level size :965, group counts :979 are the reference's lines too), with
 *     orc_sample_min behind texture(): every mip of orc_hiz has the same bits at ten resolutions up to 8K, odd and degenerate ones
 *     included, and the shader's `>` bound check only produces stores Vulkan discards;
 *   - the data layouts (include/vkv_abi.h static_asserts == the reference headers compiled as C++), packVisBuffer / unpackVisBuffer,
 *     the resolve pass's fromLinear (srgb.h.glsl compiled against glm: bit-identical);
 *   - camera.cpp's reverseDepth / generateCameraFrustum, glm perspective / lookAt, fastgltf::math node matrices, fastgltf's
 *     convertComponent; meshoptimizer's codecs, scan partition, buildMeshlets / optimizeMeshlet / computeMeshletBounds.
 *
 * Arithmetic policy, compiled with -ffp-contract=off and no fast-math:
 *   fp32 round-to-nearest, no FMA contraction;
 *   dot(a,b)      = ((a.x*b.x + a.y*b.y) + a.z*b.z) [+ a.w*b.w]
 *   mat4*vec4     = ((c0*x + c1*y) + c2*z) + c3*w per component; mat4*mat4 column by column
 *   mat3*vec3     = (c0*x + c1*y) + c2*z
 *   det(mat3)     = dot(c0, cross-like(c1,c2)) (expansion along the first column)
 *   min(x,y) = y<x?y:x ; max(x,y) = x<y?y:x ; clamp = min(max(x,lo),hi)   (GLSL spec definitions)
 *   floor(log2(x)) = exact binary exponent
 *
 * All pointers are HOST pointers; the u64 "device address" fields of vkv_Primitive and
 * vkv_VisbufferPushConstants carry host addresses here.
 */
#ifndef VKV_ORACLE_H
#define VKV_ORACLE_H

#include "../include/vkv_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* per-draw status written by orc_cull */
enum {
	ORC_FRUSTUM_CULLED = 0,
	ORC_OCCLUDED = 1,
	ORC_VISIBLE = 2,
	ORC_NOT_TESTED = 3,
	ORC_STATUS_MASK = 3,
	/* diagnostic flags (ambiguous = a comparison within a few ulp of its threshold; excluded-and-counted class) */
	ORC_AMBIG_FRUSTUM = 1 << 2,
	ORC_AMBIG_HIZ = 1 << 3,
	ORC_AMBIG_LEVEL = 1 << 4,   /* max(w,h) within rounding noise of a power of two */
	ORC_CROSSES_CAMERA = 1 << 5, /* some AABB corner has clip.w <= 0 */
	ORC_AMBIG_FOOTPRINT = 1 << 6, /* sample position within rounding noise of a texel-footprint change (or frac within 1e-4 of 0) */
	ORC_CONE_CULLED = 1 << 7      /* orc_cull_cone only: rejected by the normal cone (class ORC_FRUSTUM_CULLED) == VKV_ST_CONE_CULLED */
};

typedef struct orc_counters {
	uint64_t tested, frustum_culled, occluded, visible;
	uint64_t ambig_frustum, ambig_hiz, ambig_level, ambig_footprint, crosses_camera;
	/* raster */
	uint64_t meshlets, triangles_in, triangles_culled_facing, triangles_rejected, triangles_clipped,
	         triangles_degenerate, triangles_rasterised, fragments, fragments_passed, tie_pixels;
} orc_counters;

/* The ORC_AMBIG_* flags say "a different but equally valid evaluation of the same GLSL could decide this draw differently":
 * every threshold comparison of the task shader is checked against the first-order rounding noise of its inputs (8 roundings of
 * the largest term of each mat4*vec4 sum, propagated through the divide).  tests/test_oracle.py holds the oracle to that
 * definition against the reference's shader text compiled with glm, which associates those sums differently.
 * on = 0 skips the bookkeeping (the timed CPU baseline); decisions are unaffected. */
void orc_set_diagnostics(int on);

/* Pyramid storage: one contiguous float array, mip k at offsets[k], extent w[k] x h[k].
 * Returns mipLevels (application.cpp:472-473); total floats in *total. */
uint32_t orc_pyramid_layout(uint32_t W, uint32_t H, uint32_t offsets[17], uint32_t w[16], uint32_t h[16], uint32_t* total);

/* visbuffer.task.glsl:25-76 + culling.h.glsl:8-56 + sampler state application.cpp:438-453.
 * vp_select: 0 = camera.prevOcclusionViewProjection (reference behaviour; pass A)
 *            1 = camera.viewProjection (two-pass extension: pass B re-test against the current pyramid)
 * only_status: if non-NULL, only draws with (only_status[i]&3)==ORC_OCCLUDED are tested (pass B); others get ORC_NOT_TESTED.
 * status[N] receives the per-draw result; threads<=0 -> hardware_concurrency. */
int orc_cull(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const float* pyramid,
             int vp_select, const uint8_t* only_status, uint8_t* status, orc_counters* ctr, int threads);

/* orc_cull plus the optional normal-cone stage between the frustum and the occlusion test (extension; the reference has no cone
 * cull): cone_table[primitiveIndex] = host address of that primitive's vkv_MeshletCone[] (vkvh_scene_host_cones); NULL = orc_cull.
 * Pass B (only_status != NULL) never needs it: its inputs passed the cone test in pass A. */
int orc_cull_cone(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const float* pyramid,
                  int vp_select, const uint8_t* only_status, uint8_t* status, orc_counters* ctr, int threads, const uint64_t* cone_table);

/* visbuffer.mesh.glsl:30-104 + fixed-function state (application.cpp:326-340,772-841;
 * pipeline_builder.cpp:225-277) + visbuffer.frag.glsl:36.
 * Rasterises the given MeshletDraw indices IN ORDER onto (depth, ids_ref, ids_min, tie) which are in/out
 * (clear first with orc_clear).  ids_ref follows the reference tie rule (>= : last writer wins),
 * ids_min keeps the lowest id among exact-depth ties (what a 64-bit atomicMin produces), tie[p]=1 where
 * the final depth is shared by >= 2 different ids. */
int orc_raster(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H,
               const uint32_t* draw_ids, uint32_t n_draws,
               float* depth, uint32_t* ids_ref, uint32_t* ids_min, uint8_t* tie,
               orc_counters* ctr, int threads);

void orc_clear(uint32_t W, uint32_t H, float* depth, uint32_t* ids_ref, uint32_t* ids_min, uint8_t* tie);

/* hiz_reduce.comp.glsl:21-31 + application.cpp:951-1003 + min-sampler footprint rule.
 * Mips whose dispatch size is 0 in either axis are left untouched. */
int orc_hiz(uint32_t W, uint32_t H, const float* depth, float* pyramid, int threads);

/* The sampler itself, exposed for unit tests: min over the <=2x2 non-zero-weight texels around (u,v)
 * of a w x h image with CLAMP_TO_EDGE.  *ambig is OR-ed with 1 if a frac is within 1e-4 of 0. */
float orc_sample_min(const float* img, uint32_t w, uint32_t h, float u, float v, int* ambig);
float orc_from_linear(float c); /* srgb.h.glsl:26-32, one channel */

/* The mesh shader's per-vertex / per-triangle results for n MeshletDraws (visbuffer.mesh.glsl:43-44 mvp, :61 gl_Position, :71 transformDet,
 * :86-102 gl_CullPrimitiveEXT), laid out per draw as clip[64][4], cull[126] (0 / 1, 0xff beyond triangleCount), det[126], tdet, ambig[126],
 * noise[126] (noise = first-order bound on how far an equally valid evaluation of the same GLSL can move det; ambig = |det| <= noise, or
 * transformDet within its own noise: such an evaluation may decide the triangle differently).  tests/test_oracle.py holds this against the reference's own lines compiled with glm. */
void orc_mesh_shader(const vkv_VisbufferPushConstants* pc, const uint32_t* draw_ids, uint32_t n, float* clip, uint8_t* cull, float* det,
                     float* tdet, uint8_t* ambig, float* noise);

/* shaders/visbuffer/visbuffer_resolve.comp.glsl:17-41 + srgb.h.glsl:26-32 + dispatch application.cpp:943
 * (renderResolution.x / 32 groups of 32 threads: columns >= (W/32)*32 are never touched).  ids = the R32_UINT visbuffer;
 * out = W*H RGBA8 texels (R in the low byte), in/out: untouched pixels keep their contents.  pow() is libm powf here and
 * CUDA powf on the device: the 8-bit result may differ by one code in rare cases — tests allow +-1 per channel. */
int orc_resolve(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const uint32_t* ids, uint32_t* out);

/* visbuffer.frag.glsl:38 + visbuffer.mesh.glsl:44-45,61-63: the motion-vector attachment (R16G16_SFLOAT, cleared to 0; application.cpp:250-267,
 * 786-799), evaluated per pixel for the triangle `ids` names.  out_f: 2 floats per pixel before the fp16 store (may be NULL), out_h: the
 * attachment's 2 halves per pixel (may be NULL).  PARITY: the interpolation arithmetic is implementation-defined in Vulkan; the formula here is
 * the definition the CUDA pass is held to bit for bit, and llvmpipe running the reference's own line 38 agrees within fp32 noise. */
int orc_motion_vectors(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const uint32_t* ids, float* out_f, uint16_t* out_h);
uint16_t orc_to_half(float f);

/* 64-bit visbuffer key: (~floatBits(depth) << 32) | id */
uint64_t orc_vis64_key(float depth, uint32_t id);

#ifdef __cplusplus
}
#endif
#endif
