/* accessors.cpp — CPU restatement of the accessor conversions the reference applies while it fills glsl::Vertex and the index
 * vector (src/vk_gltf_viewer/assets.cpp:308-320): fastgltf::iterateAccessor<glm::vec3> on POSITION and
 * fastgltf::copyFromAccessor<std::uint32_t> on the indices, i.e. per component fastgltf::internal::convertComponent<float, T>
 * (submodules/fastgltf/include/fastgltf/tools.hpp:266-289): float(x), or for normalized integers
 * max(float(x) / float(numeric_limits<T>::max()), -1) (KHR_mesh_quantization), and plain widening for indices.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY PINNED: tests/test_accessors.py compares every one of the 2 x (2^8 + 2^8 +
 * 2^16 + 2^16) possible inputs against fastgltf's own function compiled from the reference tree (oracle/_ref/libref_shim.so)
 * and against the SHA-256 of those tables frozen in tests/golden/accessor_tables.json.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

extern "C" {

/* glTF componentType: 5120 BYTE, 5121 UNSIGNED_BYTE, 5122 SHORT, 5123 UNSIGNED_SHORT, 5125 UNSIGNED_INT, 5126 FLOAT */
float orc_convert_component(int type, int normalized, int value) {
	float f, mx;
	switch (type) {
	case 5120: f = (float)(int8_t)value; mx = 127.0f; break;
	case 5121: f = (float)(uint8_t)value; mx = 255.0f; break;
	case 5122: f = (float)(int16_t)value; mx = 32767.0f; break;
	case 5123: f = (float)(uint16_t)value; mx = 65535.0f; break;
	default: return 0.0f;
	}
	if (!normalized) return f;
	f = f / mx;
	return f < -1.0f ? -1.0f : f; /* fastgltf::max(x, -1): (a > b) ? a : b */
}

static size_t component_size(int type) { return type == 5120 || type == 5121 ? 1 : type == 5122 || type == 5123 ? 2 : 4; }

/* POSITION accessor (VEC3) -> glsl::Vertex[count] (24 bytes each; everything but position zero, as the reference leaves it) */
int orc_assemble_vertices(const void* src, int type, int normalized, size_t byte_stride, size_t count, void* vertices24) {
	const size_t cs = component_size(type);
	if (type != 5126 && type != 5120 && type != 5121 && type != 5122 && type != 5123) return -1;
	if (byte_stride == 0) byte_stride = 3 * cs;
	memset(vertices24, 0, count * 24);
	for (size_t i = 0; i < count; ++i) {
		const unsigned char* p = (const unsigned char*)src + i * byte_stride;
		float* out = (float*)((unsigned char*)vertices24 + i * 24);
		for (int k = 0; k < 3; ++k) {
			if (type == 5126) { memcpy(&out[k], p + 4 * k, 4); continue; }
			int v;
			if (cs == 1) v = type == 5120 ? (int)(int8_t)p[k] : (int)p[k];
			else { uint16_t u; memcpy(&u, p + 2 * k, 2); v = type == 5122 ? (int)(int16_t)u : (int)u; }
			out[k] = orc_convert_component(type, normalized, v);
		}
	}
	return 0;
}

/* index accessor (SCALAR u8 / u16 / u32) -> u32[count] */
int orc_widen_indices(const void* src, int type, size_t count, uint32_t* out) {
	if (type != 5121 && type != 5123 && type != 5125) return -1;
	for (size_t i = 0; i < count; ++i) {
		if (type == 5121) out[i] = ((const uint8_t*)src)[i];
		else if (type == 5123) { uint16_t u; memcpy(&u, (const unsigned char*)src + 2 * i, 2); out[i] = u; }
		else memcpy(&out[i], (const unsigned char*)src + 4 * i, 4);
	}
	return 0;
}

} // extern "C"
