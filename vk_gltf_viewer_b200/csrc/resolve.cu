// resolve.cu — visibility-buffer resolve: 64-bit visbuffer -> RGBA8 colour image.
// Replaces shaders/visbuffer/visbuffer_resolve.comp.glsl:17-41 (+ srgb.h.glsl:26-32, unpackVisBuffer visbuffer.h.glsl:62-65)
// and its dispatch (application.cpp:917-949).  SURVEY §8f row f1: the direct consumer of the hot path's output.
//
// The resolved colour depends on the MATERIAL only (fromLinear(albedoFactor)), so the sRGB transfer function is evaluated
// once per material by a prologue kernel; the per-pixel kernel is integer work: key -> drawIndex -> MeshletDraw.primitiveIndex
// -> Primitive.materialIndex -> colour table.  One thread per two pixels (16-byte key loads, 8-byte colour stores):
// HBM bound, 8 B read + 4 B written per pixel.
//
// Reference quirk kept: the dispatch covers (W / 32) * 32 columns only (application.cpp:943: renderResolution.x / 32
// groups of 32 threads), pixels right of that are never touched — not even cleared.
#include "kernels.cuh"

namespace {

// srgb.h.glsl:26-32 per channel, then the RGBA8_UNORM store conversion (round to nearest of clamp(v,0,1) * 255)
__device__ __forceinline__ float from_linear(float c) {
	const bool cutoff = c < 0.0031308f;
	const float higher = 1.055f * powf(c, 1.f / 2.4f) - 0.055f;
	const float lower = c * 12.92f;
	return cutoff ? lower : higher; // mix(higher, lower, cutoff)
}
__device__ __forceinline__ uint32_t unorm8(float v) {
	v = (v > 0.0f) ? v : 0.0f; // NaN -> 0
	v = (v < 1.0f) ? v : 1.0f;
	return (uint32_t)__float2int_rn(v * 255.0f);
}

__global__ void material_colors_kernel(const vkv_Material* __restrict__ materials, uint32_t n, uint32_t* __restrict__ out) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const float4 a = __ldg((const float4*)materials[i].albedoFactor);
		out[i] = unorm8(from_linear(a.x)) | (unorm8(from_linear(a.y)) << 8) | (unorm8(from_linear(a.z)) << 16) | (unorm8(a.w) << 24);
	}
}

__device__ __forceinline__ uint32_t resolve_one(const ResolveParams& p, unsigned long long key) {
	const uint32_t id = (uint32_t)key;
	if (id == VKV_VISBUFFER_CLEAR) return 0u;                 // comp.glsl:25-29: cleared to vec4(0), nothing drawn
	const uint32_t drawIndex = id >> VKV_TRIANGLE_BITS;       // unpackVisBuffer
	const uint32_t prim = __ldg(&p.draws[drawIndex].primitiveIndex);
	const uint32_t mat = __ldg(&p.primitives[prim].materialIndex);
	return __ldg(p.matColors + mat);
}

__global__ void __launch_bounds__(256) resolve_kernel(const ResolveParams p) {
	const uint32_t covered = (p.W / 32u) * 32u; // columns the reference dispatch reaches (even: pairs never straddle it)
	const uint32_t pairsPerRow = covered >> 1;
	const size_t nPairs = (size_t)pairsPerRow * p.H;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nPairs; i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t y = (uint32_t)(i / pairsPerRow), x = (uint32_t)(i % pairsPerRow) * 2u;
		const size_t px = (size_t)y * p.W + x;
		uint32_t c0, c1;
		if ((px & 1) == 0) { // 16-byte aligned pair
			const ulonglong2 k = __ldcs((const ulonglong2*)(p.vis + px));
			c0 = resolve_one(p, k.x); c1 = resolve_one(p, k.y);
			*(uint2*)(p.color + px) = make_uint2(c0, c1);
		} else {             // odd W: rows alternate alignment
			c0 = resolve_one(p, p.vis[px]); c1 = resolve_one(p, p.vis[px + 1]);
			p.color[px] = c0; p.color[px + 1] = c1;
		}
	}
}

} // namespace

cudaError_t launch_material_colors(const vkv_Material* materials, uint32_t n, uint32_t* out, int num_sms, cudaStream_t stream) {
	if (!n) return cudaSuccess;
	uint32_t grid = (n + 127) / 128;
	if (grid > (uint32_t)num_sms * 8) grid = (uint32_t)num_sms * 8;
	material_colors_kernel<<<grid, 128, 0, stream>>>(materials, n, out);
	return cudaGetLastError();
}

cudaError_t launch_resolve(const ResolveParams& p, int num_sms, cudaStream_t stream) {
	const size_t nPairs = (size_t)((p.W / 32u) * 16u) * p.H;
	if (!nPairs) return cudaSuccess;
	size_t grid = (nPairs + 255) / 256;
	if (grid > (size_t)num_sms * 16) grid = (size_t)num_sms * 16;
	resolve_kernel<<<(unsigned)grid, 256, 0, stream>>>(p);
	return cudaGetLastError();
}
