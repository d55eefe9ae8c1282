// ref_shim.cpp — extern "C" window onto the parts of the REFERENCE that compile as C++ (built by oracle/build_ref.sh
// into oracle/_ref/libref_shim.so).  TEST INFRASTRUCTURE ONLY: used to pin the CPU oracle and the host input
// generators against the reference's own code.  Everything included below is read in place from /root/reference.
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>

#include <fastgltf/tools.hpp>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>

#include "mesh_common.h.glsl"          // reference: glsl::Camera, Meshlet, Vertex, MeshletDraw, Primitive, Material
#include "visbuffer/visbuffer.h.glsl"  // reference: glsl::VisbufferPushConstants, packVisBuffer
#include "culling_head.h.glsl"         // reference: culling.h.glsl:1-30 (generated slice, see build_ref.sh)
#include "culling_tail.h.glsl"         // reference: culling.h.glsl:31-56 aabbPositions + projectAabb (generated slice, array syntax rewritten)
#include "task_lines.inc"              // reference: visbuffer.task.glsl:57-61 mip selection + sample position (generated slice)
#include "mesh_lines.inc"              // reference: visbuffer.mesh.glsl:44,61,65,71,90-98 vertex transform, determinants, facing decision (generated slice)
#include "camera_fns.inc"              // reference: camera.cpp reverseDepth + generateCameraFrustum
#include "srgb_lines.h.glsl"           // reference: srgb.h.glsl fromLinear / toLinear (generated: the `.rgb` swizzles rewritten)

// ---- stand-ins for what hiz_reduce.comp.glsl reads besides its own text: the bindless heaps (resource_table.h.glsl:12-17), the compute
// built-in, texture() and imageStore().  The sampler behind texture() is a callback (the oracle's orc_sample_min: the Vulkan min-reduction
// sampler is fixed function, there is no reference code for it); imageStore drops out-of-bounds texels as Vulkan does ("Texel Output
// Validation": an invalid texel coordinate makes the write have no effect) and counts them.
#include <fastgltf/util.hpp>
namespace fg = fastgltf;
typedef float (*ref_sample_fn)(const float* img, uint32_t w, uint32_t h, float u, float v, int* ambig);
namespace glsl {
struct RefSampledImage { const float* img; uint32_t w, h; };
struct RefStorageImage { float* img; uint32_t w, h; };
static thread_local RefSampledImage sampled_textures_heap[2];
static thread_local RefStorageImage writeonly_image2d_r32f_heap[2];
static thread_local uvec3 gl_GlobalInvocationID;
static thread_local ref_sample_fn hizSampler;
static thread_local uint64_t hizDroppedStores;
inline vec4 texture(const RefSampledImage& s, vec2 uv) { return vec4(hizSampler(s.img, s.w, s.h, uv.x, uv.y, nullptr), 0.f, 0.f, 1.f); }
inline void imageStore(const RefStorageImage& im, ivec2 p, vec4 v) {
	if (p.x < 0 || p.y < 0 || (uint32_t)p.x >= im.w || (uint32_t)p.y >= im.h) { ++hizDroppedStores; return; }
	im.img[(size_t)p.y * im.w + p.x] = v.x;
}
} // namespace glsl
#include "hiz_lines.inc"               // reference: hiz_reduce.comp.glsl:11-15,21-31 push constants + main (generated slice)
#include "hiz_dispatch.inc"            // reference: application.cpp:472-473,965,979 mip count, level size, group counts (generated slice)

#include <atomic>
#include <cmath>
#include <thread>
#include <vector>

#include <fastgltf/math.hpp>

extern "C" {

// The whole "HiZ reduction" zone (application.cpp:951-1003) with the reference's own shader text: for every dispatch the loop at :964-979
// records, every invocation of every 32x32 workgroup runs hiz_reduce.comp.glsl's main.  View 0 is the depth image, view i the pyramid's
// mip i-1 (application.cpp:503-529); the pyramid's mips live where the caller says (offset / extent per mip: Vulkan's max(1, base >> k)
// of the (W>>1, H>>1) image, application.cpp:482-487).  Returns the reference's mip count (application.cpp:472-473); dropped[0] = stores
// Vulkan discards (the shader's bound check is `>` where `>=` was meant: invocations at pos == imageSize run and write out of bounds).
uint32_t ref_hiz_reduce(uint32_t W, uint32_t H, const float* depth, float* pyramid, const uint32_t* mipOff, const uint32_t* mipW, const uint32_t* mipH,
                        uint32_t levels, ref_sample_fn sample, uint64_t* dropped) {
	const glm::u32vec2 renderResolution(W, H);
	const uint32_t mipLevels = hizMipLevels(renderResolution);
	glsl::hizSampler = sample;
	glsl::hizDroppedStores = 0;
	for (std::uint32_t i = 1; i < mipLevels + 1 && i <= levels; ++i) { // depthPyramidViews.size() == mipLevels + 1
		glm::u32vec2 levelSize, groups;
		hizDispatch(renderResolution, i, levelSize, groups);
		glsl::sampled_textures_heap[0] = (i == 1) ? glsl::RefSampledImage{depth, W, H} : glsl::RefSampledImage{pyramid + mipOff[i - 2], mipW[i - 2], mipH[i - 2]};
		glsl::writeonly_image2d_r32f_heap[1] = glsl::RefStorageImage{pyramid + mipOff[i - 1], mipW[i - 1], mipH[i - 1]};
		glsl::pushConstants.sourceImage = 0;
		glsl::pushConstants.outputImage = 1;
		glsl::pushConstants.imageSize = levelSize;
		for (uint32_t gy = 0; gy < groups.y; ++gy)
			for (uint32_t gx = 0; gx < groups.x; ++gx)
				for (uint32_t ly = 0; ly < 32; ++ly)      // layout(local_size_x = 32, local_size_y = 32), comp.glsl:9
					for (uint32_t lx = 0; lx < 32; ++lx) {
						glsl::gl_GlobalInvocationID = glm::uvec3(gx * 32 + lx, gy * 32 + ly, 0);
						glsl::hizReduceMain();
					}
	}
	if (dropped) *dropped = glsl::hizDroppedStores;
	return mipLevels;
}

// culling.h.glsl:8-19
int ref_is_aabb_in_frustum(const float c[3], const float e[3], const float frustum[24]) {
	glm::vec4 f[6];
	for (int i = 0; i < 6; ++i) f[i] = glm::vec4(frustum[i * 4], frustum[i * 4 + 1], frustum[i * 4 + 2], frustum[i * 4 + 3]);
	return glsl::isAabbInFrustum(glm::vec3(c[0], c[1], c[2]), glm::vec3(e[0], e[1], e[2]), f) ? 1 : 0;
}

// culling.h.glsl:22-29
void ref_world_aabb_extent(const float e[3], const float m[16], float out[3]) {
	glm::mat4 t;
	std::memcpy(&t, m, 64);
	glm::vec3 r = glsl::getWorldSpaceAabbExtent(glm::vec3(e[0], e[1], e[2]), t);
	out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// task.glsl:50 as glm evaluates it on the C++ side (NOTE: glm's mat4*vec4 pairs the adds, the oracle's policy is
// left-to-right; the test compares within 1 ulp and reports how many differ)
void ref_transform_point(const float m[16], const float p[3], float out[3]) {
	glm::mat4 t;
	std::memcpy(&t, m, 64);
	glm::vec4 r = t * glm::vec4(p[0], p[1], p[2], 1.0f);
	out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// culling.h.glsl:44-56 alone
void ref_project_aabb(const float c[3], const float e[3], const float vp[16], float out[6]) {
	glm::mat4 m;
	std::memcpy(&m, vp, 64);
	auto r = glsl::projectAabb(glm::vec3(c[0], c[1], c[2]), glm::vec3(e[0], e[1], e[2]), m);
	out[0] = r[0].x; out[1] = r[0].y; out[2] = r[0].z; out[3] = r[1].x; out[4] = r[1].y; out[5] = r[1].z;
}

// The task shader's per-draw decision (visbuffer.task.glsl:44-64) evaluated with the REFERENCE's code as glm evaluates it on the
// C++ side: transform * vec4(center, 1), getWorldSpaceAabbExtent, isAabbInFrustum, projectAabb and the mip-selection lines are
// the reference's own text.  Only the texture fetch is not reference code: `sample` is the oracle's min-sampler
// (orc_sample_min), and the lod clamp follows the sampler state (application.cpp:451-452: lod in [0,16], then existing mips;
// NaN -> 0).  status[i] = 0 frustum-culled, 1 occluded, 2 visible — the oracle's ORC_* values.
// NOTE glm pairs the adds of mat4*vec4 as (c0x + c1y) + (c2z + c3w); the oracle's policy is left to right.  The two evaluations
// may therefore differ in the last bits, which is exactly what the oracle's ORC_AMBIG_* / ORC_CROSSES_CAMERA flags are for.
void ref_task_cull(const glsl::VisbufferPushConstants* pc, const float* pyramid, const uint32_t* mipOff, const uint32_t* mipW, const uint32_t* mipH,
                   uint32_t levels, int vp_select, ref_sample_fn sample, uint8_t* status, int threads) {
	const auto* draws = reinterpret_cast<const glsl::MeshletDraw*>(pc->drawBuffer);
	const auto* transforms = reinterpret_cast<const glm::mat4*>(pc->transformBuffer);
	const auto* prims = reinterpret_cast<const glsl::Primitive*>(pc->primitiveBuffer);
	const glsl::Camera camera = *reinterpret_cast<const glsl::Camera*>(pc->cameraBuffer);
	const glm::ivec2 pyramidSize((int)mipW[0], (int)mipH[0]);
	const uint32_t N = pc->meshletDrawCount;
	std::atomic<uint32_t> next{0};
	auto work = [&]() {
		for (;;) {
			const uint32_t b = next.fetch_add(4096);
			if (b >= N) break;
			const uint32_t e = b + 4096 < N ? b + 4096 : N;
			for (uint32_t i = b; i < e; ++i) {
				const glsl::MeshletDraw draw = draws[i];
				const glm::mat4 transformMatrix = transforms[draw.transformIndex];
				const glsl::Primitive& primitive = prims[draw.primitiveIndex];
				const glsl::Meshlet meshlet = reinterpret_cast<const glsl::Meshlet*>(primitive.meshletBuffer)[draw.meshletIndex];
				const glm::vec3 worldAabbCenter = glm::vec3(transformMatrix * glm::vec4(meshlet.aabbCenter, 1.0f));              // task.glsl:50
				const glm::vec3 worldAabbExtent = glsl::getWorldSpaceAabbExtent(meshlet.aabbExtents, transformMatrix);          // :51
				glm::vec4 fr[6];
				for (int k = 0; k < 6; ++k) fr[k] = camera.frustum[k];
				bool visible = glsl::isAabbInFrustum(worldAabbCenter, worldAabbExtent, fr);                                      // :52
				if (!visible) { status[i] = 0; continue; }
				const auto projectedAabb = glsl::projectAabb(worldAabbCenter, worldAabbExtent,
				                                             vp_select ? camera.viewProjection : camera.prevOcclusionViewProjection);      // :56
				float level;
				glm::vec2 center;
				glsl::taskMipAndCenter(projectedAabb, pyramidSize, level, center);                                                 // :57-61
				int lod = 0;
				if (level == level) { const float cl = level < 0.f ? 0.f : (level > 16.f ? 16.f : level); lod = (int)cl; }
				if (lod > (int)levels - 1) lod = (int)levels - 1;
				const float depth = sample(pyramid + mipOff[lod], mipW[lod], mipH[lod], center.x, center.y, nullptr);          // :62
				visible = visible && depth < projectedAabb[1].z;                                                                   // :64
				status[i] = visible ? 2 : 1;
			}
		}
	};
	int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
	if (nt < 1) nt = 1;
	std::vector<std::thread> pool;
	for (int t = 1; t < nt; ++t) pool.emplace_back(work);
	work();
	for (auto& t : pool) t.join();
}

// The mesh shader's arithmetic (visbuffer.mesh.glsl:43-44 mvp, :61 clip position, :65 clipVertices, :71 transformDet, :86-98 the facing
// decision) evaluated with the REFERENCE's own lines as glm evaluates them on the C++ side, for the given MeshletDraws.  Per draw d:
// clip[d][v] = gl_Position of meshlet vertex v (4 floats, maxVertices = 64 slots), cull[d][t] = gl_CullPrimitiveEXT of triangle t
// (126 slots; 0 / 1, 0xff = slot beyond triangleCount), det[d][t] = the determinant the decision was taken on, tdet[d] = transformDet.
// NOTE glm's mat4*mat4, mat4*vec4 and determinant associate differently from the oracle's stated policy; the test allows the
// decisions to differ only where the oracle flags the determinant as within rounding noise of zero.
void ref_mesh_shader(const glsl::VisbufferPushConstants* pc, const uint32_t* draw_ids, uint32_t n, float* clip, uint8_t* cull, float* det, float* tdet) {
	const auto* draws = reinterpret_cast<const glsl::MeshletDraw*>(pc->drawBuffer);
	const auto* transforms = reinterpret_cast<const glm::mat4*>(pc->transformBuffer);
	const auto* prims = reinterpret_cast<const glsl::Primitive*>(pc->primitiveBuffer);
	const auto* materials = reinterpret_cast<const glsl::Material*>(pc->materialBuffer);
	const glsl::Camera camera = *reinterpret_cast<const glsl::Camera*>(pc->cameraBuffer);
	for (uint32_t d = 0; d < n; ++d) {
		const glsl::MeshletDraw draw = draws[draw_ids[d]];                                                                       // :31
		const glsl::Primitive& primitive = prims[draw.primitiveIndex];                                                           // :33
		const glsl::Meshlet meshlet = reinterpret_cast<const glsl::Meshlet*>(primitive.meshletBuffer)[draw.meshletIndex];        // :34
		const glsl::Material& material = materials[primitive.materialIndex];                                                     // :35
		const glm::mat4 transformMatrix = transforms[draw.transformIndex];                                                       // :43
		const glm::mat4 mvp = glsl::meshMvp(camera, transformMatrix);                                                            // :44
		const auto* vertexIndices = reinterpret_cast<const uint32_t*>(primitive.vertexIndexBuffer);
		const auto* vertices = reinterpret_cast<const glsl::Vertex*>(primitive.vertexBuffer);
		const auto* primitiveIndices = reinterpret_cast<const uint8_t*>(primitive.primitiveIndexBuffer);
		glm::vec3 clipVertices[glsl::maxVertices];
		for (uint32_t v = 0; v < meshlet.vertexCount && v < glsl::maxVertices; ++v) {
			const glsl::Vertex& vertex = vertices[vertexIndices[meshlet.vertexOffset + v]];                                      // :57-58
			const glm::vec4 pos = glsl::meshVertex(mvp, vertex, clipVertices[v]);                                                // :61,65
			std::memcpy(clip + ((size_t)d * glsl::maxVertices + v) * 4, &pos, 16);
		}
		const float transformDet = glsl::meshTransformDet(transformMatrix);                                                      // :71
		tdet[d] = transformDet;
		for (uint32_t t = 0; t < glsl::maxPrimitives; ++t) {
			uint8_t& c = cull[(size_t)d * glsl::maxPrimitives + t];
			float& dt = det[(size_t)d * glsl::maxPrimitives + t];
			c = 0xff; dt = 0.0f;
			if (t >= meshlet.triangleCount) continue;
			const glm::uvec3 indices(primitiveIndices[meshlet.triangleOffset + t * 3 + 0], primitiveIndices[meshlet.triangleOffset + t * 3 + 1],
			                         primitiveIndices[meshlet.triangleOffset + t * 3 + 2]);                                       // :77-80
			if (!material.doubleSided) c = glsl::meshCull(clipVertices, indices, transformDet, dt) ? 1 : 0;                       // :86-98
			else c = 0;                                                                                                          // :99-102
		}
	}
}

// visbuffer_resolve.comp.glsl:39 `vec4 resolved = fromLinear(material.albedoFactor)` with the reference's srgb.h.glsl:26-32 against glm
void ref_from_linear(const float in[4], float out[4]) {
	const glm::vec4 r = glsl::fromLinear(glm::vec4(in[0], in[1], in[2], in[3]));
	out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
// visbuffer.h.glsl:62-65
void ref_unpack_visbuffer(uint32_t v, uint32_t* drawIndex, uint32_t* primitiveId) { glsl::unpackVisBuffer(v, *drawIndex, *primitiveId); }
uint32_t ref_visbuffer_clear_value() { return glsl::visbufferClearValue; }

uint32_t ref_pack_visbuffer(uint32_t drawIndex, uint32_t primitiveId) { return glsl::packVisBuffer(drawIndex, primitiveId); }

// sizeof/offsetof table of the shared structs, in the order tests/test_abi.py expects
int ref_layout(uint32_t* out, int cap) {
	const uint32_t v[] = {
		(uint32_t)sizeof(glsl::Camera), (uint32_t)offsetof(glsl::Camera, prevOcclusionViewProjection), (uint32_t)offsetof(glsl::Camera, viewProjection),
		(uint32_t)offsetof(glsl::Camera, occlusionViewProjection), (uint32_t)offsetof(glsl::Camera, frustum),
		(uint32_t)sizeof(glsl::Meshlet), (uint32_t)offsetof(glsl::Meshlet, triangleOffset), (uint32_t)offsetof(glsl::Meshlet, vertexCount),
		(uint32_t)offsetof(glsl::Meshlet, triangleCount), (uint32_t)offsetof(glsl::Meshlet, aabbExtents), (uint32_t)offsetof(glsl::Meshlet, aabbCenter),
		(uint32_t)sizeof(glsl::Vertex), (uint32_t)offsetof(glsl::Vertex, color), (uint32_t)offsetof(glsl::Vertex, normal), (uint32_t)offsetof(glsl::Vertex, uv),
		(uint32_t)sizeof(glsl::MeshletDraw), (uint32_t)offsetof(glsl::MeshletDraw, meshletIndex), (uint32_t)offsetof(glsl::MeshletDraw, transformIndex),
		(uint32_t)sizeof(glsl::Primitive), (uint32_t)offsetof(glsl::Primitive, meshletBuffer), (uint32_t)offsetof(glsl::Primitive, aabbExtents),
		(uint32_t)offsetof(glsl::Primitive, aabbCenter), (uint32_t)offsetof(glsl::Primitive, meshletCount), (uint32_t)offsetof(glsl::Primitive, materialIndex),
		(uint32_t)sizeof(glsl::Material), (uint32_t)offsetof(glsl::Material, albedoIndex), (uint32_t)offsetof(glsl::Material, uvOffset),
		(uint32_t)offsetof(glsl::Material, alphaCutoff), (uint32_t)offsetof(glsl::Material, doubleSided),
		(uint32_t)sizeof(glsl::VisbufferPushConstants), (uint32_t)offsetof(glsl::VisbufferPushConstants, meshletDrawCount),
		(uint32_t)offsetof(glsl::VisbufferPushConstants, transformBuffer), (uint32_t)offsetof(glsl::VisbufferPushConstants, primitiveBuffer),
		(uint32_t)offsetof(glsl::VisbufferPushConstants, cameraBuffer), (uint32_t)offsetof(glsl::VisbufferPushConstants, materialBuffer),
		(uint32_t)offsetof(glsl::VisbufferPushConstants, depthPyramid),
		glsl::maxVertices, glsl::maxPrimitives, glsl::maxMeshlets, glsl::triangleBits, glsl::drawIndexBits,
	};
	const int n = (int)(sizeof(v) / sizeof(v[0]));
	for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
	return n;
}

// Camera::updateCamera's matrix assembly (camera.cpp:170-193) with the reference's own glm + reverseDepth + generateCameraFrustum
void ref_camera(const float eye[3], const float center[3], const float up[3], uint32_t W, uint32_t H, float vp_out[16], float frustum_out[24]) {
	auto view = glm::lookAtRH(glm::vec3(eye[0], eye[1], eye[2]), glm::vec3(center[0], center[1], center[2]), glm::vec3(up[0], up[1], up[2]));
	static constexpr auto zNear = 0.1f;
	static constexpr auto zFar = 1000.0f;
	static constexpr auto fov = glm::radians(75.0f);
	const auto aspectRatio = static_cast<float>(W) / static_cast<float>(H);
	auto projectionMatrix = glm::perspectiveRH_ZO(fov, aspectRatio, zNear, zFar);
	projectionMatrix[1][1] *= -1;
	glm::mat4 vp = reverseDepth(projectionMatrix) * view;
	std::memcpy(vp_out, &vp, 64);
	std::array<glm::vec4, 6> fr;
	generateCameraFrustum(vp, fr);
	std::memcpy(frustum_out, fr.data(), 96);
}

// world.cpp:221 : scale(rotate(translate(parent, T), R), S) with fastgltf::math
void ref_node_matrix(const float parent[16], const float t[3], const float r[4], const float s[3], float out[16]) {
	namespace fm = fastgltf::math;
	fm::fmat4x4 p;
	std::memcpy(p.data(), parent, 64);
	auto m = fm::scale(fm::rotate(fm::translate(p, fm::fvec3(t[0], t[1], t[2])), fm::fquat(r[0], r[1], r[2], r[3])), fm::fvec3(s[0], s[1], s[2]));
	std::memcpy(out, m.data(), 64);
}

// fastgltf::math::decomposeTransformMatrix (math.hpp:854-891): what Options::DecomposeNodeMatrices applies to node.matrix
void ref_decompose(const float m[16], float t[3], float r[4], float s[3]) {
	namespace fm = fastgltf::math;
	fm::fmat4x4 mat;
	std::memcpy(mat.data(), m, 64);
	fm::fvec3 scale, translation;
	fm::fquat rotation;
	fm::decomposeTransformMatrix(mat, scale, rotation, translation);
	for (int i = 0; i < 3; ++i) { t[i] = translation[i]; s[i] = scale[i]; }
	for (int i = 0; i < 4; ++i) r[i] = rotation[i];
}

// fastgltf::internal::convertComponent<float, T> (tools.hpp:266-289): what iterateAccessor<glm::vec3> applies to every POSITION
// component (assets.cpp:310-314).  type: glTF componentType (5120 BYTE, 5121 UNSIGNED_BYTE, 5122 SHORT, 5123 UNSIGNED_SHORT)
float ref_convert_component(int type, int normalized, int value) {
	switch (type) {
	case 5120: return fastgltf::internal::convertComponent<float, std::int8_t>((std::int8_t)value, normalized != 0);
	case 5121: return fastgltf::internal::convertComponent<float, std::uint8_t>((std::uint8_t)value, normalized != 0);
	case 5122: return fastgltf::internal::convertComponent<float, std::int16_t>((std::int16_t)value, normalized != 0);
	case 5123: return fastgltf::internal::convertComponent<float, std::uint16_t>((std::uint16_t)value, normalized != 0);
	default: return 0.0f;
	}
}

} // extern "C"
