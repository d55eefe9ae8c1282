// scene.hpp — host-side scene container: what World + AssetLoadTask hold in the reference
// (world.hpp / assets.hpp), reduced to the buffers the geometry path consumes.
#pragma once
#include "../../include/vkv_host.h"
#include "hmath.hpp"

#include <string>
#include <vector>

namespace vkvh {

struct MeshletRec { uint32_t vertex_offset, triangle_offset, vertex_count, triangle_count; };

struct PrimitiveData {
	std::vector<vkv_Vertex> vertices;
	std::vector<uint32_t> meshletVertices;   // Primitive.vertexIndexBuffer
	std::vector<uint8_t> meshletTriangles;   // Primitive.primitiveIndexBuffer
	std::vector<vkv_Meshlet> meshlets;       // Primitive.meshletBuffer
	std::vector<int16_t> qpos;               // extension: the accessor's original SHORT positions, 4 per vertex (x, y, z, 0), when the asset
	bool qnormalized = false;                // is KHR_mesh_quantization'd — the side buffer of the in-register dequantisation (vkv_set_quantized_positions)
	std::vector<vkv_MeshletCone> cones;      // extension: normal cone per meshlet (meshopt_computeMeshletBounds), side buffer of the cone cull
	vkv_Primitive header{};                  // addresses filled per address space
	uint64_t triangles = 0;
};

struct Node {
	int32_t parent = -1;
	bool hasMesh = false;                    // takes a transform slot (world.cpp:246-252), even when its mesh has no drawable primitive
	std::vector<int32_t> primitives;         // the node's mesh: every primitive shares the node's transform (world.cpp:253-262)
	float t[3] = {0, 0, 0}, r[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
	std::vector<int32_t> children;
};

void build_meshlets_builtin(const std::vector<vkv_Vertex>& vertices, const std::vector<uint32_t>& indices,
                            std::vector<MeshletRec>& meshlets, std::vector<uint32_t>& meshletVertices,
                            std::vector<uint8_t>& meshletTriangles);

bool builder_is_injected(); // vkvh_set_meshlet_builder installed foreign entry points (they are then called from one thread only)
bool build_primitive(PrimitiveData& pd, std::vector<vkv_Vertex>&& vertices, const uint32_t* indices, uint32_t index_count, uint32_t material_index);
std::vector<vkv_Vertex> vertices_from_positions(const float* positions, uint32_t vertex_count);

} // namespace vkvh

struct vkvh_scene;
namespace vkvh { int32_t add_built_primitive(vkvh_scene* s, PrimitiveData&& pd); }

struct vkvh_scene {
	std::vector<vkvh::PrimitiveData> primitives;
	std::vector<vkv_Material> materials;
	std::vector<vkvh::Node> nodes;
	std::vector<int32_t> roots;
	std::vector<vkv_MeshletDraw> draws;
	std::vector<float> transforms; // 16 per mesh node, traversal order
	std::vector<vkv_Primitive> hostPrimitives;
	std::vector<std::vector<vkv_MeshletCone>> hostCones; // per primitive, material-adjusted (double-sided -> disabled)
	std::vector<uint64_t> hostConeTable;
	bool finalized = false;
	// default-view hints set by the procedural generators
	int kind = 0; // 0 custom, 1 icosphere, 2 atrium, 3 lattice, 4 city
	float boundsMin[3] = {0, 0, 0}, boundsMax[3] = {0, 0, 0};
};
