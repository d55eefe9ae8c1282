// accessors.cu — glTF accessor -> renderer buffer conversions on the device.
// Replaces the two host loops of PrimitiveProcessingTask::processPrimitive that fill glsl::Vertex and the u32 index vector
// (src/vk_gltf_viewer/assets.cpp:308-320): fastgltf::iterateAccessor<glm::vec3> on POSITION and
// fastgltf::copyFromAccessor<std::uint32_t> on the indices.  Per component that is fastgltf::internal::convertComponent<float, T>
// (submodules/fastgltf/include/fastgltf/tools.hpp:266-289): float(x), or — KHR_mesh_quantization's normalized integers —
// max(float(x) / float(numeric_limits<T>::max()), -1) with an IEEE division.  Everything of a Vertex but `position` is zero,
// as the reference leaves it.  HBM-bound element-wise kernels: one thread per vertex (one 16-byte + one 8-byte store) / per index.
#include "kernels.cuh"

namespace {

__device__ __forceinline__ float convert_component(int type, bool normalized, int v) {
	float f, mx;
	switch (type) {
	case 5120: f = (float)(signed char)v; mx = 127.0f; break;
	case 5121: f = (float)(unsigned char)v; mx = 255.0f; break;
	case 5122: f = (float)(short)v; mx = 32767.0f; break;
	default: f = (float)(unsigned short)v; mx = 65535.0f; break;
	}
	if (!normalized) return f;
	f = f / mx;                       // IEEE division (the file is compiled without fast-math)
	return f < -1.0f ? -1.0f : f;     // fastgltf::max(x, -1)
}

__global__ void assemble_vertices_kernel(const uint8_t* __restrict__ src, int type, int normalized, uint32_t stride, uint32_t count,
                                         uint32_t* __restrict__ out /* 6 words per vertex */) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		const uint8_t* p = src + (size_t)i * stride;
		float v[3];
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			if (type == 5126) v[k] = __uint_as_float((uint32_t)p[4 * k] | (uint32_t)p[4 * k + 1] << 8 | (uint32_t)p[4 * k + 2] << 16 | (uint32_t)p[4 * k + 3] << 24);
			else if (type == 5120 || type == 5121) v[k] = convert_component(type, normalized != 0, p[k]);
			else v[k] = convert_component(type, normalized != 0, (int)((uint32_t)p[2 * k] | (uint32_t)p[2 * k + 1] << 8));
		}
		uint32_t* o = out + (size_t)i * 6; // 24-byte records: 8-byte aligned, not 16
		*(uint2*)o = make_uint2(__float_as_uint(v[0]), __float_as_uint(v[1]));
		*(uint2*)(o + 2) = make_uint2(__float_as_uint(v[2]), 0u);
		*(uint2*)(o + 4) = make_uint2(0u, 0u);
	}
}

__global__ void widen_indices_kernel(const uint8_t* __restrict__ src, int type, uint32_t count, uint32_t* __restrict__ out) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		uint32_t v;
		if (type == 5121) v = src[i];
		else if (type == 5123) v = (uint32_t)src[2 * (size_t)i] | (uint32_t)src[2 * (size_t)i + 1] << 8;
		else v = (uint32_t)src[4 * (size_t)i] | (uint32_t)src[4 * (size_t)i + 1] << 8 | (uint32_t)src[4 * (size_t)i + 2] << 16 | (uint32_t)src[4 * (size_t)i + 3] << 24;
		out[i] = v;
	}
}

uint32_t grid_for(uint32_t n, int num_sms) {
	uint32_t g = (n + 255) / 256;
	const uint32_t cap = (uint32_t)num_sms * 16;
	return g > cap ? cap : (g ? g : 1);
}

} // namespace

cudaError_t launch_assemble_vertices(const uint8_t* src, int type, int normalized, uint32_t stride, uint32_t count, void* vertices, int num_sms, cudaStream_t stream) {
	if (count) assemble_vertices_kernel<<<grid_for(count, num_sms), 256, 0, stream>>>(src, type, normalized, stride, count, (uint32_t*)vertices);
	return cudaGetLastError();
}
cudaError_t launch_widen_indices(const uint8_t* src, int type, uint32_t count, uint32_t* out, int num_sms, cudaStream_t stream) {
	if (count) widen_indices_kernel<<<grid_for(count, num_sms), 256, 0, stream>>>(src, type, count, out);
	return cudaGetLastError();
}
