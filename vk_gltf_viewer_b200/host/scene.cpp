// scene.cpp — host-side input generators mirroring the reference's asset / world / camera code (see vkv_host.h).
#include "scene.hpp"

#include <algorithm>
#include <cassert>
#include <functional>

using namespace vkvh;

namespace {
vkvh_build_bound_fn g_bound = nullptr;
vkvh_build_fn g_build = nullptr;
vkvh_optimize_fn g_optimize = nullptr;
bool g_morton = false;
} // namespace

extern "C" {

void vkvh_set_meshlet_builder(vkvh_build_bound_fn bound, vkvh_build_fn build, vkvh_optimize_fn optimize) {
	g_bound = bound; g_build = build; g_optimize = optimize;
}

void vkvh_select_builder(int morton) { g_morton = morton != 0; }

vkvh_scene* vkvh_scene_new(void) {
	auto* s = new vkvh_scene();
	// assets.cpp:486-492 : default material at index 0
	vkv_Material m{};
	m.albedoFactor[0] = m.albedoFactor[1] = m.albedoFactor[2] = m.albedoFactor[3] = 1.0f;
	m.albedoIndex = 0xFFFFFFFFu; // invalidHandle
	m.uvScale[0] = m.uvScale[1] = 1.0f;
	m.alphaCutoff = 0.5f;
	m.doubleSided = 0;
	s->materials.push_back(m);
	return s;
}

void vkvh_scene_free(vkvh_scene* s) { delete s; }

uint32_t vkvh_scene_add_material(vkvh_scene* s, const float albedo[4], int double_sided) {
	vkv_Material m = s->materials[0];
	for (int i = 0; i < 4; ++i) m.albedoFactor[i] = albedo[i];
	m.doubleSided = double_sided ? 1u : 0u;
	s->materials.push_back(m);
	return (uint32_t)s->materials.size() - 1; // glTF material i -> i+1 (assets.cpp:292-294)
}

} // extern "C"

// processPrimitive (assets.cpp:288-373) for one primitive: meshlets + bounds.  Touches no scene state, so the procedural
// generators may run it for many primitives side by side (the reference does the same through its task scheduler).
bool vkvh::builder_is_injected() { return g_build && g_bound; }

bool vkvh::build_primitive(PrimitiveData& pd, std::vector<vkv_Vertex>&& vertices, const uint32_t* indices, uint32_t index_count, uint32_t material_index) {
	if (vertices.empty() || index_count < 3) return false;
	for (uint32_t i = 0; i < index_count; ++i)
		if (indices[i] >= vertices.size()) return false;
	pd.vertices = std::move(vertices);
	pd.triangles = index_count / 3;
	std::vector<uint32_t> idx(indices, indices + (index_count / 3) * 3);

	std::vector<MeshletRec> recs;
	const bool injected = g_build && g_bound;
	if (injected || !g_morton) {
		// assets.cpp:322-346 — through the built-in restatement of meshoptimizer's builder (host/clusterizer.cpp) or the injected
		// meshoptimizer-compatible entry points
		const vkvh_build_bound_fn boundFn = injected ? g_bound : vkvh_meshlets_bound;
		const vkvh_build_fn buildFn = injected ? g_build : vkvh_meshlets_build;
		const vkvh_optimize_fn optimizeFn = injected ? g_optimize : vkvh_meshlet_optimize;
		const size_t maxTris = VKV_MAX_MESHLET_TRIANGLES;
		size_t bound = boundFn(idx.size(), VKV_MAX_VERTICES, maxTris);
		std::vector<vkvh_meshopt_Meshlet> ms(bound);
		pd.meshletVertices.resize(bound * VKV_MAX_VERTICES);
		pd.meshletTriangles.resize(bound * maxTris * 3);
		size_t n = buildFn(ms.data(), pd.meshletVertices.data(), pd.meshletTriangles.data(), idx.data(), idx.size(),
		                   pd.vertices[0].position, pd.vertices.size(), sizeof(vkv_Vertex), VKV_MAX_VERTICES, maxTris, 0.0f);
		if (n == 0) return false;
		const auto& last = ms[n - 1];
		pd.meshletVertices.resize(last.vertex_count + last.vertex_offset);
		pd.meshletTriangles.resize(((last.triangle_count * 3 + 3) & ~3u) + last.triangle_offset);
		ms.resize(n);
		for (auto& m : ms) {
			if (optimizeFn) optimizeFn(&pd.meshletVertices[m.vertex_offset], &pd.meshletTriangles[m.triangle_offset], m.triangle_count, m.vertex_count);
			recs.push_back({m.vertex_offset, m.triangle_offset, m.vertex_count, m.triangle_count});
		}
	} else {
		build_meshlets_builtin(pd.vertices, idx, recs, pd.meshletVertices, pd.meshletTriangles);
	}

	// per-meshlet bounds: assets.cpp:349-372
	pd.meshlets.reserve(recs.size());
	float pmin[3] = {1e30f, 1e30f, 1e30f}, pmax[3] = {-1e30f, -1e30f, -1e30f};
	for (const auto& r : recs) {
		const float* p0 = pd.vertices[pd.meshletVertices[r.vertex_offset]].position;
		float mn[3] = {p0[0], p0[1], p0[2]}, mx[3] = {p0[0], p0[1], p0[2]};
		for (uint32_t i = 1; i < r.vertex_count; ++i) {
			const float* p = pd.vertices[pd.meshletVertices[r.vertex_offset + i]].position;
			for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); }
		}
		vkv_Meshlet m{};
		m.vertexOffset = r.vertex_offset;
		m.triangleOffset = r.triangle_offset;
		m.vertexCount = (uint8_t)r.vertex_count;
		m.triangleCount = (uint8_t)r.triangle_count;
		for (int k = 0; k < 3; ++k) {
			float c = (mn[k] + mx[k]) * 0.5f;
			m.aabbCenter[k] = c;
			m.aabbExtents[k] = mx[k] - c;
			pmin[k] = std::min(pmin[k], mn[k]); pmax[k] = std::max(pmax[k], mx[k]);
		}
		pd.meshlets.push_back(m);
		// extension: the meshlet's normal cone (meshopt_computeMeshletBounds; the reference never computes it, assets.cpp:323)
		vkvh_meshopt_Bounds b;
		vkvh_meshlet_bounds(&pd.meshletVertices[r.vertex_offset], &pd.meshletTriangles[r.triangle_offset], r.triangle_count, pd.vertices[0].position,
		                    pd.vertices.size(), sizeof(vkv_Vertex), &b);
		vkv_MeshletCone cn{};
		for (int k = 0; k < 3; ++k) { cn.apex[k] = b.cone_apex[k]; cn.axis[k] = b.cone_axis[k]; }
		cn.cutoff = b.cone_cutoff >= 1.0f ? 2.0f : b.cone_cutoff; // 1 = "normals wider than a hemisphere": never reject
		pd.cones.push_back(cn);
	}
	// assets.cpp:303-306 (accessor min/max -> primitive AABB)
	for (int k = 0; k < 3; ++k) {
		pd.header.aabbCenter[k] = (pmin[k] + pmax[k]) / 2.0f;
		pd.header.aabbExtents[k] = pmax[k] - pd.header.aabbCenter[k];
	}
	pd.header.meshletCount = (uint32_t)pd.meshlets.size();
	pd.header.materialIndex = material_index;
	return true;
}

int32_t vkvh::add_built_primitive(vkvh_scene* s, PrimitiveData&& pd) {
	if (pd.header.materialIndex >= s->materials.size()) return -1;
	s->primitives.push_back(std::move(pd));
	s->finalized = false;
	return (int32_t)s->primitives.size() - 1;
}

static int32_t add_primitive_vertices(vkvh_scene* s, std::vector<vkv_Vertex>&& vertices, const uint32_t* indices,
                                      uint32_t index_count, uint32_t material_index) {
	if (material_index >= s->materials.size()) return -1;
	PrimitiveData pd;
	if (!build_primitive(pd, std::move(vertices), indices, index_count, material_index)) return -1;
	return add_built_primitive(s, std::move(pd));
}

extern "C" {

} // extern "C"
std::vector<vkv_Vertex> vkvh::vertices_from_positions(const float* positions, uint32_t vertex_count) {
	std::vector<vkv_Vertex> v(vertex_count);
	for (uint32_t i = 0; i < vertex_count; ++i) {
		std::memset(&v[i], 0, sizeof(vkv_Vertex));
		v[i].position[0] = positions[i * 3]; v[i].position[1] = positions[i * 3 + 1]; v[i].position[2] = positions[i * 3 + 2];
		v[i].color[0] = v[i].color[1] = v[i].color[2] = v[i].color[3] = 255;
	}
	return v;
}
extern "C" {

int32_t vkvh_scene_add_primitive(vkvh_scene* s, const float* positions, uint32_t vertex_count, const uint32_t* indices,
                                 uint32_t index_count, uint32_t material_index) {
	return add_primitive_vertices(s, vertices_from_positions(positions, vertex_count), indices, index_count, material_index);
}

int32_t vkvh_scene_add_primitive_i16(vkvh_scene* s, const int16_t* positions, uint32_t vertex_count, int normalized,
                                     const uint32_t* indices, uint32_t index_count, uint32_t material_index) {
	std::vector<vkv_Vertex> v(vertex_count);
	for (uint32_t i = 0; i < vertex_count; ++i) {
		std::memset(&v[i], 0, sizeof(vkv_Vertex));
		for (int k = 0; k < 3; ++k) {
			// fastgltf tools.hpp:266-289 convertComponent: float(x), or max(float(x)/32767, -1) when normalized
			float f = (float)positions[i * 3 + k];
			if (normalized) f = std::max(f / 32767.0f, -1.0f);
			v[i].position[k] = f;
		}
		v[i].color[0] = v[i].color[1] = v[i].color[2] = v[i].color[3] = 255;
	}
	const int32_t idx = add_primitive_vertices(s, std::move(v), indices, index_count, material_index);
	if (idx >= 0) { // keep the accessor's own 16-bit data: the rasteriser can read it instead of the expanded 24-byte Vertex records
		auto& pd = s->primitives[idx];
		pd.qpos.resize((size_t)vertex_count * 4);
		for (uint32_t i = 0; i < vertex_count; ++i) {
			pd.qpos[i * 4 + 0] = positions[i * 3 + 0]; pd.qpos[i * 4 + 1] = positions[i * 3 + 1]; pd.qpos[i * 4 + 2] = positions[i * 3 + 2]; pd.qpos[i * 4 + 3] = 0;
		}
		pd.qnormalized = normalized != 0;
	}
	return idx;
}

int vkvh_scene_upload_quantized(vkvh_scene* s, vkvh_upload_fn upload, void* user, uint64_t* table_addr) {
	if (!s || !upload || !table_addr) return -1;
	std::vector<vkv_QuantizedPositions> table(s->primitives.size());
	for (size_t i = 0; i < s->primitives.size(); ++i) {
		const auto& p = s->primitives[i];
		table[i] = vkv_QuantizedPositions{0, 0, 0};
		if (p.qpos.empty()) continue;
		int rc = upload(user, p.qpos.data(), p.qpos.size() * 2, &table[i].positions);
		if (rc) return rc;
		table[i].normalized = p.qnormalized ? 1u : 0u;
	}
	return upload(user, table.data(), table.size() * sizeof(vkv_QuantizedPositions), table_addr);
}

int32_t vkvh_scene_add_node_trs(vkvh_scene* s, int32_t parent, int32_t primitive, const float t[3], const float r[4], const float sc[3]) {
	return vkvh_scene_add_node_mesh(s, parent, primitive >= 0 ? &primitive : nullptr, primitive >= 0 ? 1u : 0u, primitive >= 0 ? 1 : 0, t, r, sc);
}

int32_t vkvh_scene_add_node_mesh(vkvh_scene* s, int32_t parent, const int32_t* primitives, uint32_t n_primitives, int has_mesh, const float t[3],
                                 const float r[4], const float sc[3]) {
	if (parent >= (int32_t)s->nodes.size() || (n_primitives && !primitives)) return -1;
	for (uint32_t i = 0; i < n_primitives; ++i)
		if (primitives[i] < 0 || primitives[i] >= (int32_t)s->primitives.size()) return -1;
	Node n;
	n.parent = parent; n.hasMesh = has_mesh != 0 || n_primitives > 0;
	n.primitives.assign(primitives, primitives + n_primitives);
	if (t) std::memcpy(n.t, t, 12);
	if (r) std::memcpy(n.r, r, 16);
	if (sc) std::memcpy(n.s, sc, 12);
	s->nodes.push_back(n);
	int32_t idx = (int32_t)s->nodes.size() - 1;
	if (parent >= 0) s->nodes[parent].children.push_back(idx); else s->roots.push_back(idx);
	s->finalized = false;
	return idx;
}

int vkvh_scene_finalize(vkvh_scene* s) {
	s->draws.clear(); s->transforms.clear();
	uint32_t transformCount = 0;
	// world.cpp:187-228 iterateNode ; :242-264 draws ; :308-319 transforms
	std::function<void(int32_t, const mat4&)> walk = [&](int32_t ni, const mat4& parent) {
		const Node& n = s->nodes[ni];
		mat4 m = scale(rotate(translate(parent, n.t), n.r), n.s);
		if (n.hasMesh) {
			uint32_t ti = transformCount++;
			for (int32_t prim : n.primitives) {
				const auto& pd = s->primitives[prim];
				for (uint32_t i = 0; i < pd.header.meshletCount; ++i) s->draws.push_back(vkv_MeshletDraw{(uint32_t)prim, i, ti});
			}
			s->transforms.insert(s->transforms.end(), m.m, m.m + 16);
		}
		for (int32_t c : n.children) walk(c, m);
	};
	for (int32_t r : s->roots) walk(r, identity());
	if (s->draws.size() > VKV_MAX_MESHLET_DRAWS) return -2; // 25-bit drawIndex (visbuffer.h.glsl:15-17)
	s->hostPrimitives.clear();
	for (auto& pd : s->primitives) {
		pd.header.vertexIndexBuffer = (uint64_t)(uintptr_t)pd.meshletVertices.data();
		pd.header.primitiveIndexBuffer = (uint64_t)(uintptr_t)pd.meshletTriangles.data();
		pd.header.vertexBuffer = (uint64_t)(uintptr_t)pd.vertices.data();
		pd.header.meshletBuffer = (uint64_t)(uintptr_t)pd.meshlets.data();
		s->hostPrimitives.push_back(pd.header);
	}
	s->finalized = true;
	return 0;
}

void vkvh_scene_counts(const vkvh_scene* s, vkvh_counts* c) {
	std::memset(c, 0, sizeof(*c));
	c->primitives = (uint32_t)s->primitives.size();
	c->materials = (uint32_t)s->materials.size();
	c->transforms = (uint32_t)(s->transforms.size() / 16);
	c->draws = (uint32_t)s->draws.size();
	c->nodes = (uint32_t)s->nodes.size();
	for (const auto& p : s->primitives) {
		c->triangles_unique += p.triangles; c->meshlets_unique += p.meshlets.size(); c->vertices_unique += p.vertices.size();
	}
	for (const auto& d : s->draws) c->triangles_instanced += s->primitives[d.primitiveIndex].meshlets[d.meshletIndex].triangleCount;
}

const vkv_MeshletDraw* vkvh_scene_draws(const vkvh_scene* s) { return s->draws.data(); }
const float* vkvh_scene_transforms(const vkvh_scene* s) { return s->transforms.data(); }
const vkv_Material* vkvh_scene_materials(const vkvh_scene* s) { return s->materials.data(); }

int vkvh_scene_primitive(const vkvh_scene* s, uint32_t index, vkvh_primitive_view* out) {
	if (index >= s->primitives.size()) return -1;
	const auto& p = s->primitives[index];
	out->vertex_indices = p.meshletVertices.data(); out->vertex_indices_count = p.meshletVertices.size();
	out->triangles = p.meshletTriangles.data(); out->triangles_bytes = p.meshletTriangles.size();
	out->vertices = p.vertices.data(); out->vertex_count = p.vertices.size();
	out->meshlets = p.meshlets.data(); out->meshlet_count = p.meshlets.size();
	out->header = p.header;
	return 0;
}

int vkvh_scene_host_pc(vkvh_scene* s, const vkv_Camera* camera, vkv_VisbufferPushConstants* out) {
	if (!s->finalized && vkvh_scene_finalize(s) != 0) return -1;
	std::memset(out, 0, sizeof(*out));
	out->drawBuffer = (uint64_t)(uintptr_t)s->draws.data();
	out->meshletDrawCount = (uint32_t)s->draws.size();
	out->transformBuffer = (uint64_t)(uintptr_t)s->transforms.data();
	out->primitiveBuffer = (uint64_t)(uintptr_t)s->hostPrimitives.data();
	out->cameraBuffer = (uint64_t)(uintptr_t)camera;
	out->materialBuffer = (uint64_t)(uintptr_t)s->materials.data();
	return 0;
}

static void material_adjusted_cones(vkvh_scene* s) {
	s->hostCones.assign(s->primitives.size(), {});
	s->hostConeTable.assign(s->primitives.size(), 0);
	for (size_t i = 0; i < s->primitives.size(); ++i) {
		const auto& p = s->primitives[i];
		s->hostCones[i] = p.cones;
		if (s->materials[p.header.materialIndex].doubleSided)
			for (auto& c : s->hostCones[i]) c.cutoff = 2.0f; // mesh.glsl:86: no facing cull for double-sided materials
		s->hostConeTable[i] = (uint64_t)(uintptr_t)s->hostCones[i].data();
	}
}

int vkvh_scene_host_cones(vkvh_scene* s, const uint64_t** table) {
	if (!s || !table) return -1;
	material_adjusted_cones(s);
	*table = s->hostConeTable.data();
	return 0;
}

int vkvh_scene_upload_cones(vkvh_scene* s, vkvh_upload_fn upload, void* user, uint64_t* table_addr) {
	if (!s || !upload || !table_addr) return -1;
	material_adjusted_cones(s);
	// one allocation for all cones, one for the table
	std::vector<vkv_MeshletCone> all;
	std::vector<uint64_t> first(s->primitives.size());
	for (size_t i = 0; i < s->primitives.size(); ++i) { first[i] = all.size(); all.insert(all.end(), s->hostCones[i].begin(), s->hostCones[i].end()); }
	uint64_t base = 0;
	int rc = upload(user, all.data(), all.size() * sizeof(vkv_MeshletCone), &base);
	if (rc) return rc;
	std::vector<uint64_t> table(s->primitives.size());
	for (size_t i = 0; i < table.size(); ++i) table[i] = base + first[i] * sizeof(vkv_MeshletCone);
	return upload(user, table.data(), table.size() * 8, table_addr);
}

int vkvh_scene_upload(vkvh_scene* s, vkvh_upload_fn upload, void* user, const vkv_Camera* camera, vkv_VisbufferPushConstants* out) {
	if (!s->finalized && vkvh_scene_finalize(s) != 0) return -1;
	std::memset(out, 0, sizeof(*out));
	std::vector<vkv_Primitive> dev(s->primitives.size());
	int rc = 0;
	for (size_t i = 0; i < s->primitives.size() && rc == 0; ++i) {
		const auto& p = s->primitives[i];
		dev[i] = p.header;
		// assets.cpp:377-424 : four buffers per primitive
		rc = upload(user, p.meshletVertices.data(), p.meshletVertices.size() * 4, &dev[i].vertexIndexBuffer);
		if (!rc) rc = upload(user, p.meshletTriangles.data(), p.meshletTriangles.size(), &dev[i].primitiveIndexBuffer);
		if (!rc) rc = upload(user, p.vertices.data(), p.vertices.size() * sizeof(vkv_Vertex), &dev[i].vertexBuffer);
		if (!rc) rc = upload(user, p.meshlets.data(), p.meshlets.size() * sizeof(vkv_Meshlet), &dev[i].meshletBuffer);
	}
	if (!rc) rc = upload(user, dev.data(), dev.size() * sizeof(vkv_Primitive), &out->primitiveBuffer);             // world.cpp:89-130
	if (!rc) rc = upload(user, s->materials.data(), s->materials.size() * sizeof(vkv_Material), &out->materialBuffer); // world.cpp:132-178
	if (!rc) rc = upload(user, s->draws.data(), s->draws.size() * sizeof(vkv_MeshletDraw), &out->drawBuffer);      // world.cpp:267-290
	if (!rc) rc = upload(user, s->transforms.data(), s->transforms.size() * 4, &out->transformBuffer);             // world.cpp:321-344
	if (!rc) rc = upload(user, camera, sizeof(vkv_Camera), &out->cameraBuffer);                                    // camera.cpp:86-104
	out->meshletDrawCount = (uint32_t)s->draws.size();
	return rc;
}

// camera.cpp:70-84
void vkvh_frustum_from_vp(const float vp[16], float frustum[6][4]) {
	for (int i = 0; i < 4; ++i) { frustum[0][i] = vp[i * 4 + 3] + vp[i * 4 + 0]; }
	for (int i = 0; i < 4; ++i) { frustum[1][i] = vp[i * 4 + 3] - vp[i * 4 + 0]; }
	for (int i = 0; i < 4; ++i) { frustum[2][i] = vp[i * 4 + 3] + vp[i * 4 + 1]; }
	for (int i = 0; i < 4; ++i) { frustum[3][i] = vp[i * 4 + 3] - vp[i * 4 + 1]; }
	for (int i = 0; i < 4; ++i) { frustum[4][i] = vp[i * 4 + 3] + vp[i * 4 + 2]; }
	for (int i = 0; i < 4; ++i) { frustum[5][i] = vp[i * 4 + 3] - vp[i * 4 + 2]; }
	for (int p = 0; p < 6; ++p) {
		float* pl = frustum[p];
		float len = std::sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]); // glm::length(vec3) = sqrt(dot)
		for (int i = 0; i < 4; ++i) pl[i] = pl[i] / len;
		pl[3] = -pl[3];
	}
}

// camera.cpp:170-193
void vkvh_camera_update(vkv_Camera* cam, const float eye[3], const float center[3], const float up[3], uint32_t width,
                        uint32_t height, int first) {
	const mat4 view = lookAtRH({eye[0], eye[1], eye[2]}, {center[0], center[1], center[2]}, {up[0], up[1], up[2]});
	const float zNear = 0.1f, zFar = 1000.0f;
	const float fov = 75.0f * 0.01745329251994329576923690768489f; // glm::radians(75.f)
	const float aspect = (float)width / (float)height;
	mat4 proj = perspectiveRH_ZO(fov, aspect, zNear, zFar);
	proj.m[5] *= -1.0f;
	// camera.cpp:38-48 reverseDepth
	mat4 reverseZ = identity();
	reverseZ.m[10] = -1.0f; reverseZ.m[14] = 1.0f;
	const mat4 vp = mul(mul(reverseZ, proj), view);
	if (first) {
		std::memcpy(cam->prevViewProjection, vp.m, 64);
		std::memcpy(cam->prevOcclusionViewProjection, vp.m, 64);
	} else {
		std::memcpy(cam->prevViewProjection, cam->viewProjection, 64);
		std::memcpy(cam->prevOcclusionViewProjection, cam->occlusionViewProjection, 64);
	}
	std::memcpy(cam->viewProjection, vp.m, 64);
	std::memcpy(cam->occlusionViewProjection, vp.m, 64);
	vkvh_frustum_from_vp(cam->viewProjection, cam->frustum);
}

} // extern "C"
