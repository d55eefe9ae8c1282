/* vkv_abi.h — binary data contract of the per-frame geometry hot path.
 *
 * Plain-C restatement of the scalar-block-layout structs the reference shares
 * between its GLSL shaders and its C++ host code.  Every struct below is what
 * the reference *uploads*; the CUDA path consumes these bytes unchanged, so a
 * maintainer can hand the very same host arrays to vkv_upload().
 *
 *   reference (relative to the upstream tree)           here
 *   shaders/mesh_common.h.glsl:20-27   Camera           vkv_Camera
 *   shaders/mesh_common.h.glsl:49-58   Meshlet          vkv_Meshlet
 *   shaders/mesh_common.h.glsl:60-71   Vertex           vkv_Vertex
 *   shaders/mesh_common.h.glsl:96-100  MeshletDraw      vkv_MeshletDraw
 *   shaders/mesh_common.h.glsl:102-113 Primitive        vkv_Primitive
 *   shaders/mesh_common.h.glsl:115-126 Material         vkv_Material
 *   shaders/visbuffer/visbuffer.h.glsl:37-47            vkv_VisbufferPushConstants
 *   shaders/visbuffer/visbuffer.h.glsl:15-16,58-67      VKV_TRIANGLE_BITS / pack
 *   shaders/mesh_common.h.glsl:36-38   limits           VKV_MAX_*
 *
 * Matrices are column-major float[16] (m[col*4+row]), as glm::mat4.
 * BUFFER_REF fields are 64-bit device addresses (here: CUDA device pointers
 * returned by vkv_upload), exactly where the reference stores VkDeviceAddress.
 */
#ifndef VKV_ABI_H
#define VKV_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define VKV_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define VKV_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* mesh_common.h.glsl:36-38 ; assets.cpp:324 uses alignDown(126,4)=124 triangles */
#define VKV_MAX_VERTICES 64u
#define VKV_MAX_PRIMITIVES 126u
#define VKV_MAX_MESHLET_TRIANGLES 124u
#define VKV_MAX_MESHLETS_PER_TASK 102u
/* visbuffer.h.glsl:15-16 */
#define VKV_TRIANGLE_BITS 7u
#define VKV_DRAW_INDEX_BITS 25u
#define VKV_MAX_MESHLET_DRAWS (1u << VKV_DRAW_INDEX_BITS)
#define VKV_VISBUFFER_CLEAR 0xFFFFFFFFu               /* visbuffer.h.glsl:67, application.cpp:782 */
#define VKV_VIS64_CLEAR 0xFFFFFFFFFFFFFFFFull         /* == depth 0.0 (far, reverse-Z) | id clear */

typedef struct vkv_Camera {
	float prevViewProjection[16];          /* @0   */
	float prevOcclusionViewProjection[16]; /* @64  */
	float viewProjection[16];              /* @128 */
	float occlusionViewProjection[16];     /* @192 */
	float frustum[6][4];                   /* @256 : xyz = unit normal, w = NEGATED plane constant (camera.cpp:80-83) */
} vkv_Camera;

typedef struct vkv_Meshlet {
	uint32_t vertexOffset;   /* @0  into the primitive's vertex-index buffer */
	uint32_t triangleOffset; /* @4  into the primitive's u8 triangle buffer (4-byte aligned) */
	uint8_t vertexCount;     /* @8  */
	uint8_t triangleCount;   /* @9  */
	uint8_t _pad[2];
	float aabbExtents[3];    /* @12 */
	float aabbCenter[3];     /* @24 */
} vkv_Meshlet;

typedef struct vkv_Vertex {
	float position[3]; /* @0  — the only field the ID path reads */
	uint8_t color[4];  /* @12 */
	uint8_t normal[3]; /* @16 */
	uint8_t _pad;
	uint16_t uv[2];    /* @20 f16 bits */
} vkv_Vertex;

typedef struct vkv_MeshletDraw {
	uint32_t primitiveIndex;
	uint32_t meshletIndex;
	uint32_t transformIndex;
} vkv_MeshletDraw;

typedef struct vkv_Primitive {
	uint64_t vertexIndexBuffer;    /* @0  u32[]            */
	uint64_t primitiveIndexBuffer; /* @8  u8[] 3/triangle  */
	uint64_t vertexBuffer;         /* @16 vkv_Vertex[]     */
	uint64_t meshletBuffer;        /* @24 vkv_Meshlet[]    */
	float aabbExtents[3];          /* @32 */
	float aabbCenter[3];           /* @44 */
	uint32_t meshletCount;         /* @56 */
	uint32_t materialIndex;        /* @60 : 0 = default, glTF material i -> i+1 (assets.cpp:292-294) */
} vkv_Primitive;

typedef struct vkv_Material {
	float albedoFactor[4]; /* @0  */
	uint32_t albedoIndex;  /* @16 */
	float uvOffset[2];     /* @20 */
	float uvScale[2];      /* @28 */
	float uvRotation;      /* @36 */
	float alphaCutoff;     /* @40 */
	uint32_t doubleSided;  /* @44 : GLSL bool, 4 bytes — the only field the ID path reads (mesh.glsl:86) */
} vkv_Material;

typedef struct vkv_VisbufferPushConstants {
	uint64_t drawBuffer;       /* @0  vkv_MeshletDraw[] */
	uint32_t meshletDrawCount; /* @8  */
	uint32_t _pad0;
	uint64_t transformBuffer;  /* @16 float[16][]       */
	uint64_t primitiveBuffer;  /* @24 vkv_Primitive[]   */
	uint64_t cameraBuffer;     /* @32 vkv_Camera        */
	uint64_t materialBuffer;   /* @40 vkv_Material[]    */
	uint32_t depthPyramid;     /* @48 bindless handle in the reference; ignored here (context owns the pyramid) */
	uint32_t _pad1;
} vkv_VisbufferPushConstants;

/* Not a reference struct: one entry per (mesh-node, primitive) in World::rebuildDrawBuffer's traversal order
 * (world.cpp:243-266) — what vkv_build_draws expands into MeshletDraw[] on the device. */
typedef struct vkv_DrawSegment {
	uint32_t primitiveIndex;
	uint32_t transformIndex;
} vkv_DrawSegment;

VKV_STATIC_ASSERT(sizeof(vkv_Camera) == 352, "Camera");
VKV_STATIC_ASSERT(offsetof(vkv_Camera, frustum) == 256, "Camera.frustum");
VKV_STATIC_ASSERT(sizeof(vkv_Meshlet) == 36, "Meshlet");
VKV_STATIC_ASSERT(offsetof(vkv_Meshlet, vertexCount) == 8, "Meshlet.vertexCount");
VKV_STATIC_ASSERT(offsetof(vkv_Meshlet, aabbExtents) == 12, "Meshlet.aabbExtents");
VKV_STATIC_ASSERT(offsetof(vkv_Meshlet, aabbCenter) == 24, "Meshlet.aabbCenter");
VKV_STATIC_ASSERT(sizeof(vkv_Vertex) == 24, "Vertex");
VKV_STATIC_ASSERT(offsetof(vkv_Vertex, uv) == 20, "Vertex.uv");
VKV_STATIC_ASSERT(sizeof(vkv_MeshletDraw) == 12, "MeshletDraw");
VKV_STATIC_ASSERT(sizeof(vkv_Primitive) == 64, "Primitive");
VKV_STATIC_ASSERT(offsetof(vkv_Primitive, aabbExtents) == 32, "Primitive.aabbExtents");
VKV_STATIC_ASSERT(offsetof(vkv_Primitive, meshletCount) == 56, "Primitive.meshletCount");
VKV_STATIC_ASSERT(sizeof(vkv_Material) == 48, "Material");
VKV_STATIC_ASSERT(offsetof(vkv_Material, doubleSided) == 44, "Material.doubleSided");
VKV_STATIC_ASSERT(sizeof(vkv_VisbufferPushConstants) == 56, "VisbufferPushConstants");
VKV_STATIC_ASSERT(offsetof(vkv_VisbufferPushConstants, transformBuffer) == 16, "PC.transformBuffer");
VKV_STATIC_ASSERT(offsetof(vkv_VisbufferPushConstants, depthPyramid) == 48, "PC.depthPyramid");

/* ---- extension (not a reference struct): the normal cone of a meshlet, the side buffer of the optional cone cull ----------------
 * meshopt_Bounds' cone fields (meshoptimizer.h:507-541), which the reference computes nowhere (assets.cpp:323 runs the builder
 * with cone weight 0 and never calls meshopt_computeMeshletBounds).  The 36-byte Meshlet record cannot change, so the cones live
 * in a buffer of their own: one array per primitive, indexed like the primitive's Meshlet[], found through a table of device
 * addresses indexed by primitiveIndex (vkv_set_cone_table).  A meshlet is backfacing as a whole when
 *   dot(normalize(cone_apex - eye), cone_axis) >= cone_cutoff            (meshoptimizer.h:531)
 * with eye the camera position in the mesh's OWN space.  cutoff >= 2 disables the test for the meshlet (double-sided material,
 * normals spread over more than a hemisphere). */
typedef struct vkv_MeshletCone {
	float apex[3];
	float cutoff;
	float axis[3];
	float reserved;
} vkv_MeshletCone;
VKV_STATIC_ASSERT(sizeof(vkv_MeshletCone) == 32, "MeshletCone");

/* ---- extension (not a reference struct): KHR_mesh_quantization positions kept in their 16-bit form --------------------------------
 * The reference expands every POSITION accessor to f32 on the host (assets.cpp:310-314, fastgltf convertComponent) and uploads
 * 24-byte Vertex records of which the geometry path reads 12 bytes.  With this side table (one entry per primitive, indexed by
 * primitiveIndex; vkv_set_quantized_positions) the rasteriser reads 8 bytes per vertex — int16 x, y, z, 0 — and applies
 * convertComponent in registers: float(x), or max(float(x) / 32767, -1) when normalized.  Same floats, a third of the bytes.
 * positions == 0: the primitive has no 16-bit form, its Vertex buffer is read as usual. */
typedef struct vkv_QuantizedPositions {
	uint64_t positions;      /* device address of int16[4 * vertexCount] */
	uint32_t normalized;     /* accessor.normalized */
	uint32_t reserved;
} vkv_QuantizedPositions;
VKV_STATIC_ASSERT(sizeof(vkv_QuantizedPositions) == 16, "QuantizedPositions");

/* visbuffer.h.glsl:58-60 */
static inline uint32_t vkv_pack_visbuffer(uint32_t drawIndex, uint32_t primitiveId) {
	return (drawIndex << VKV_TRIANGLE_BITS) | primitiveId;
}

/* Pyramid geometry (application.cpp:472-494, 964-979).
 * mipLevels = floor(log2(max(W,H))); pyramid mip k has extent max(1,(W>>1)>>k) x max(1,(H>>1)>>k);
 * the dispatch that fills mip k writes only (W>>(k+1)) x (H>>(k+1)) texels (may be 0 => never written). */
static inline uint32_t vkv_mip_levels(uint32_t w, uint32_t h) {
	uint32_t m = w > h ? w : h, l = 0;
	while (m > 1) { m >>= 1; ++l; }
	return l;
}
static inline uint32_t vkv_mip_extent(uint32_t base, uint32_t k) { /* base = W or H of the render target */
	uint32_t e = (base >> 1) >> k;
	return e ? e : 1u;
}

#endif /* VKV_ABI_H */
