/* vkv_host.h — C ABI of the host-side input generators (libvkv_host.so).
 *
 * This is the C++ host mirror of what the reference does BEFORE the hot path: it produces, byte for byte in the
 * reference's layouts (include/vkv_abi.h), the buffers the renderer uploads.  None of it runs on the GPU and none
 * of it is timed in the frame metric.
 *
 *   reference                                                     here
 *   assets.cpp:288-373  processPrimitive (meshlets + AABBs)       vkvh_scene_add_primitive
 *   world.cpp:187-228   iterateNode  (M = scale(rotate(translate)))vkvh_scene_add_node_trs
 *   world.cpp:230-293   rebuildDrawBuffer (MeshletDraw per node x meshlet) vkvh_scene_finalize
 *   world.cpp:295-345   updateTransformBuffer                     vkvh_scene_finalize
 *   world.cpp:89-178    addAsset primitive/material upload        vkvh_scene_upload
 *   camera.cpp:38-48,70-84,170-193  updateCamera                  vkvh_camera_update
 *   assets.cpp:486-492  default material at index 0               vkvh_scene_new
 */
#ifndef VKV_HOST_H
#define VKV_HOST_H

#include "vkv_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vkvh_scene vkvh_scene;

/* Optional replacement meshlet builder with meshoptimizer's signatures (meshoptimizer.h: meshopt_buildMeshletsBound,
 * meshopt_buildMeshlets, meshopt_optimizeMeshlet).  Tests point these at the reference's own meshoptimizer build
 * (oracle/_ref) to cross-check the built-in default below; passing NULLs restores the default. */
typedef struct vkvh_meshopt_Meshlet { unsigned vertex_offset, triangle_offset, vertex_count, triangle_count; } vkvh_meshopt_Meshlet;
typedef size_t (*vkvh_build_bound_fn)(size_t index_count, size_t max_vertices, size_t max_triangles);
typedef size_t (*vkvh_build_fn)(vkvh_meshopt_Meshlet* meshlets, unsigned* meshlet_vertices, unsigned char* meshlet_triangles,
                                const unsigned* indices, size_t index_count, const float* vertex_positions, size_t vertex_count,
                                size_t vertex_positions_stride, size_t max_vertices, size_t max_triangles, float cone_weight);
typedef void (*vkvh_optimize_fn)(unsigned* meshlet_vertices, unsigned char* meshlet_triangles, size_t triangle_count, size_t vertex_count);
void vkvh_set_meshlet_builder(vkvh_build_bound_fn bound, vkvh_build_fn build, vkvh_optimize_fn optimize);

/* The built-in default (host/clusterizer.cpp): the partition the reference uploads — the algorithms of meshoptimizer 0.20's
 * meshopt_buildMeshletsBound / meshopt_buildMeshlets / meshopt_optimizeMeshlet / meshopt_computeMeshletBounds
 * (assets.cpp:322-346; submodules/meshoptimizer/src/clusterizer.cpp), restated here with the library's signatures and
 * byte-identical output (tests/test_host.py checks every byte against the reference's own build of the library).
 * vkvh_meshlets_build returns 0 for arguments the library would assert on (bad limits, an index >= vertex_count). */
typedef struct vkvh_meshopt_Bounds {
	float center[3], radius;              /* bounding sphere */
	float cone_apex[3], cone_axis[3], cone_cutoff;   /* normal cone: reject when dot(normalize(apex - eye), axis) >= cutoff (meshoptimizer.h:531) */
	signed char cone_axis_s8[3], cone_cutoff_s8;
} vkvh_meshopt_Bounds;
size_t vkvh_meshlets_bound(size_t index_count, size_t max_vertices, size_t max_triangles);
size_t vkvh_meshlets_build(vkvh_meshopt_Meshlet* meshlets, unsigned* meshlet_vertices, unsigned char* meshlet_triangles, const unsigned* indices,
                           size_t index_count, const float* vertex_positions, size_t vertex_count, size_t vertex_positions_stride,
                           size_t max_vertices, size_t max_triangles, float cone_weight);
void vkvh_meshlet_optimize(unsigned* meshlet_vertices, unsigned char* meshlet_triangles, size_t triangle_count, size_t vertex_count);
void vkvh_meshlet_bounds(const unsigned* meshlet_vertices, const unsigned char* meshlet_triangles, size_t triangle_count, const float* vertex_positions,
                         size_t vertex_count, size_t vertex_positions_stride, vkvh_meshopt_Bounds* out);
/* 0 (default): the reference's partition as above.  1: the round-1 Morton-order greedy packer (vertex-limited 64 v / ~67 t meshlets;
 * kept as a second, harder workload for A/B measurements).  Ignored while vkvh_set_meshlet_builder has injected a builder. */
void vkvh_select_builder(int morton);

/* --- scene construction ------------------------------------------------------------------------------------ */
vkvh_scene* vkvh_scene_new(void);                       /* material 0 = default (assets.cpp:486-492) */
void vkvh_scene_free(vkvh_scene*);
/* returns material index to pass to add_primitive (already +1-shifted like assets.cpp:292-294) */
uint32_t vkvh_scene_add_material(vkvh_scene*, const float albedo[4], int double_sided);
/* positions: float xyz, tightly packed.  Builds meshlets (64 v / 124 t / cone 0) and per-meshlet AABBs. Returns primitive index. */
int32_t vkvh_scene_add_primitive(vkvh_scene*, const float* positions, uint32_t vertex_count,
                                 const uint32_t* indices, uint32_t index_count, uint32_t material_index);
/* KHR_mesh_quantization: int16 non-normalised positions expanded to f32 on the host exactly as
 * fastgltf::iterateAccessor<vec3> does (tools.hpp:266-289; assets.cpp:310-314). normalized!=0 -> max(x/32767,-1). */
int32_t vkvh_scene_add_primitive_i16(vkvh_scene*, const int16_t* positions, uint32_t vertex_count, int normalized,
                                     const uint32_t* indices, uint32_t index_count, uint32_t material_index);
/* adds a mesh node with TRS (translation xyz, rotation quaternion xyzw, scale xyz); parent = -1 for a root. Returns node index.
 * primitive < 0 -> transform-only node. */
int32_t vkvh_scene_add_node_trs(vkvh_scene*, int32_t parent, int32_t primitive, const float t[3], const float r[4], const float s[3]);
/* a glTF node with a MESH: all n_primitives primitives share the node's transform slot (world.cpp:246-262); has_mesh != 0 with
 * n_primitives == 0 is a mesh without drawable primitives — it still takes a transform slot, as in the reference */
int32_t vkvh_scene_add_node_mesh(vkvh_scene*, int32_t parent, const int32_t* primitives, uint32_t n_primitives, int has_mesh,
                                 const float t[3], const float r[4], const float s[3]);
/* glTF 2.0 binary (GLB) ingest — what AssetLoadTask::loadGltf + processPrimitive + World::addAsset read of an asset
 * (assets.cpp:288-373,526-552; world.cpp:187-293): buffers (the BIN chunk, base64 data URIs), bufferViews with byteStride,
 * POSITION accessors of any component type incl. KHR_mesh_quantization (fastgltf's convertComponent rules), u8 / u16 / u32 or
 * generated indices, materials (baseColorFactor, alphaCutoff, doubleSided; index + 1), meshes with several primitives, node TRS
 * or matrices (decomposed like fastgltf::math::decomposeTransformMatrix), scenes[scene].nodes.  Returns a finalized scene or NULL
 * with a message in err (external files need vkvh_scene_load_file, which knows the asset's folder).
 * EXT_meshopt_compression (CompressedBufferDataAdapter, assets.cpp:70-171: every compressed bufferView is decoded with
 * meshopt_decodeVertexBuffer / IndexBuffer / IndexSequence, then the Oct / Quat / Exp filter): the bytes are decoded by the decoder the
 * caller installs — one call per compressed view, mode / filter numbered like vkv.h's VKV_MESHOPT_* (fastgltf's enums), dst holds
 * count * stride bytes (+ 4 of slack), return 0 or meshoptimizer's error code.  libvkv's device decoder fits (vkv_upload the stream,
 * vkv_meshopt_plan_create + vkv_meshopt_run, vkv_download); tests install the reference's meshoptimizer.  Without a decoder such assets are
 * refused with a message that says so; a stream that fails to decode is refused too (the reference ignores the code, assets.cpp:149). */
typedef int (*vkvh_meshopt_decode_fn)(void* user, uint32_t mode, uint32_t filter, uint32_t count, uint32_t stride, const void* src, size_t src_bytes, void* dst);
void vkvh_set_meshopt_decoder(vkvh_meshopt_decode_fn fn, void* user);
vkvh_scene* vkvh_scene_load_glb(const void* data, size_t bytes, char* err, size_t errcap);
/* The same for an asset FILE, .glb or .gltf (AssetLoadTask::loadGltf, assets.cpp:526-552: MappedGltfFile::FromPath + Parser::loadGltf with
 * the asset's folder) including what BufferLoadTask does (assets.cpp:36-68): buffers whose uri names a local file are read from the
 * asset's folder (percent-decoded, byteLength bytes).  Sparse accessors are densified as fastgltf's iterateAccessor reads them. */
vkvh_scene* vkvh_scene_load_file(const char* path, char* err, size_t errcap);
/* walks the node tree depth-first and emits MeshletDraw[] + transforms[] (world.cpp:230-345) */
int vkvh_scene_finalize(vkvh_scene*);

/* --- procedural scenes for the BASELINE.json configs ---------------------------------------- */
vkvh_scene* vkvh_scene_icosphere(uint32_t frequency);                                     /* cfg 1: 57 -> 64,980 tris */
vkvh_scene* vkvh_scene_atrium(uint32_t detail);                                           /* cfg 2: 128 -> 262,144 tris, int16 positions */
vkvh_scene* vkvh_scene_lattice(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t patch_quads, uint64_t seed); /* cfg 3/5 */
vkvh_scene* vkvh_scene_city(uint32_t nbx, uint32_t nby, uint32_t target_tris_per_building, uint64_t seed);  /* cfg 4 */
/* a sensible camera for view `view` of `nviews` (eye, centre) for the scene kind */
void vkvh_scene_default_view(const vkvh_scene*, uint32_t view, uint32_t nviews, float eye[3], float center[3]);

/* --- accessors ---------------------------------------------------------------------------------------------- */
typedef struct vkvh_counts {
	uint32_t primitives, materials, transforms, draws, nodes;
	uint64_t triangles_unique, triangles_instanced, meshlets_unique, vertices_unique;
} vkvh_counts;
void vkvh_scene_counts(const vkvh_scene*, vkvh_counts*);
const vkv_MeshletDraw* vkvh_scene_draws(const vkvh_scene*);
const float* vkvh_scene_transforms(const vkvh_scene*);
const vkv_Material* vkvh_scene_materials(const vkvh_scene*);
/* per primitive arrays */
typedef struct vkvh_primitive_view {
	const uint32_t* vertex_indices; uint64_t vertex_indices_count;
	const uint8_t* triangles; uint64_t triangles_bytes;
	const vkv_Vertex* vertices; uint64_t vertex_count;
	const vkv_Meshlet* meshlets; uint64_t meshlet_count;
	vkv_Primitive header;     /* buffer addresses = HOST addresses of the arrays above */
} vkvh_primitive_view;
int vkvh_scene_primitive(const vkvh_scene*, uint32_t index, vkvh_primitive_view* out);

/* Push constants over HOST memory (for the CPU oracle).  `camera` must outlive the use. */
int vkvh_scene_host_pc(vkvh_scene*, const vkv_Camera* camera, vkv_VisbufferPushConstants* out);

/* Upload everything through a callback with vkv_upload's shape and build the DEVICE push constants
 * (mirrors createMeshBuffers/addAsset/rebuildDrawBuffer/updateTransformBuffer uploads).  The camera slot is
 * uploaded once here; use vkv_update (include/vkv.h) to rewrite it per frame. */
typedef int (*vkvh_upload_fn)(void* user, const void* host, size_t bytes, uint64_t* dev_addr);
int vkvh_scene_upload(vkvh_scene*, vkvh_upload_fn upload, void* user, const vkv_Camera* camera, vkv_VisbufferPushConstants* out);

/* --- normal cones (extension: the side buffer of the optional cone cull, vkv_abi.h vkv_MeshletCone) -----------------------
 * Computed per meshlet with meshopt_computeMeshletBounds' algorithm when a primitive is added.  The table has one entry per
 * primitive: the address of its vkv_MeshletCone[meshletCount].  Primitives with a double-sided material get cutoff = 2 (never
 * rejected).  _host: HOST addresses (for the CPU oracle; valid while the scene lives); _upload: device addresses through the
 * vkv_upload-shaped callback, *table_addr then goes to vkv_set_cone_table. */
int vkvh_scene_host_cones(vkvh_scene*, const uint64_t** table);
int vkvh_scene_upload_cones(vkvh_scene*, vkvh_upload_fn upload, void* user, uint64_t* table_addr);

/* --- 16-bit positions (extension: the side buffer of the in-register dequantisation, vkv_abi.h vkv_QuantizedPositions) ------
 * Primitives added through vkvh_scene_add_primitive_i16 (or loaded from SHORT accessors) keep their 16-bit data; this uploads it
 * (8 bytes per vertex) and a per-primitive table whose device address goes to vkv_set_quantized_positions. */
int vkvh_scene_upload_quantized(vkvh_scene*, vkvh_upload_fn upload, void* user, uint64_t* table_addr);
/* cfg 4 with KHR_mesh_quantization-style geometry: every building's positions snapped to int16 units of 1/512, the node scale
 * carrying the dequantisation (what gltfpack emits); same layout and triangle counts as vkvh_scene_city */
vkvh_scene* vkvh_scene_city_quantized(uint32_t nbx, uint32_t nby, uint32_t target_tris_per_building, uint64_t seed);

/* --- camera (camera.cpp:170-193) ----------------------------------------------------------------------------- */
/* first!=0: all four matrices are set to the new viewProjection (headless start).
 * otherwise prev* <- current, then viewProjection/occlusionViewProjection/frustum are rewritten. */
void vkvh_camera_update(vkv_Camera* cam, const float eye[3], const float center[3], const float up[3],
                        uint32_t width, uint32_t height, int first);
void vkvh_frustum_from_vp(const float vp[16], float frustum[6][4]);  /* camera.cpp:70-84 */

#ifdef __cplusplus
}
#endif
#endif
