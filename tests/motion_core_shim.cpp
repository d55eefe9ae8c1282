// motion_core_shim.cpp — a HOST build of the product's motion-vector arithmetic (vk_gltf_viewer_b200/csrc/motion_core.h, the same source the
// CUDA kernel compiles) behind the same gather the kernel does, for the CPU suite: tests/test_motion.py compares it bit for bit with the
// oracle's independent restatement (orc_motion_vectors) and its fp16 conversion with numpy's.  TEST INFRASTRUCTURE: built by the test into
// oracle/_ref/ with g++ -ffp-contract=off (the host counterpart of nvcc's -fmad=false); it is not a CPU fallback — nothing in the product
// loads it, and without a GPU vkv_motion_vectors fails like every other entry point.
#include <cstddef>
#include <cstdint>

#include "vkv_abi.h"
#include "motion_core.h"

extern "C" {

// mirrors motion_kernel (motion.cu) line for line, with host addresses in the push constants
int shim_motion_vectors(const vkv_VisbufferPushConstants* pc, uint32_t W, uint32_t H, const uint32_t* ids, float* out_f, uint16_t* out_h) {
	const vkv_MeshletDraw* draws = (const vkv_MeshletDraw*)pc->drawBuffer;
	const vkv_Primitive* primitives = (const vkv_Primitive*)pc->primitiveBuffer;
	const float* transforms = (const float*)pc->transformBuffer;
	const vkv_Camera* camera = (const vkv_Camera*)pc->cameraBuffer;
	const size_t n = (size_t)W * H;
	for (size_t i = 0; i < n; ++i) {
		const uint32_t id = ids[i];
		float mv[2] = {0.f, 0.f};
		uint16_t hx = 0, hy = 0;
		if (id != VKV_VISBUFFER_CLEAR) {
			const uint32_t drawIndex = id >> VKV_TRIANGLE_BITS, tri = id & ((1u << VKV_TRIANGLE_BITS) - 1u);
			const vkv_MeshletDraw d = draws[drawIndex];
			const vkv_Primitive* prim = primitives + d.primitiveIndex;
			const vkv_Meshlet* ml = (const vkv_Meshlet*)prim->meshletBuffer + d.meshletIndex;
			const uint8_t* t3 = (const uint8_t*)prim->primitiveIndexBuffer + ml->triangleOffset + tri * 3u;
			const uint32_t* vidx = (const uint32_t*)prim->vertexIndexBuffer + ml->vertexOffset;
			const vkv_Vertex* verts = (const vkv_Vertex*)prim->vertexBuffer;
			const float* T = transforms + (size_t)d.transformIndex * 16;
			float mvp[16], prevMvp[16];
			vkv_motion::mul44m(camera->viewProjection, T, mvp);
			vkv_motion::mul44m(camera->prevViewProjection, T, prevMvp);
			vkv_motion::motion_pixel(mvp, prevMvp, verts[vidx[t3[0]]].position, verts[vidx[t3[1]]].position, verts[vidx[t3[2]]].position,
			                         (uint32_t)(i % W), (uint32_t)(i / W), W, H, mv);
			hx = vkv_motion::half_rn(mv[0]); hy = vkv_motion::half_rn(mv[1]);
		}
		if (out_f) { out_f[i * 2] = mv[0]; out_f[i * 2 + 1] = mv[1]; }
		if (out_h) { out_h[i * 2] = hx; out_h[i * 2 + 1] = hy; }
	}
	return 0;
}

void shim_half_rn(const float* in, uint16_t* out, size_t n) {
	for (size_t i = 0; i < n; ++i) out[i] = vkv_motion::half_rn(in[i]);
}

} // extern "C"
