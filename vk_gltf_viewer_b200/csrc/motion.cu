// motion.cu — the motion-vector attachment of the reference's visbuffer pass, derived from the finished visbuffer.
// Replaces visbuffer.frag.glsl:38 (+ the position / prevPosition varyings of visbuffer.mesh.glsl:44-45,61-63) and the second colour
// attachment of the pass (application.cpp:250-267 R16G16_SFLOAT "Motion vectors", cleared to 0 at :786-799; consumed at :1086-1118).
//
// The reference writes a motion vector per FRAGMENT while rasterising (overdraw included, one more 4-byte attachment in the ROP path);
// with a visibility buffer the same image is one gather per PIXEL after the depth test has settled: id -> MeshletDraw -> Meshlet ->
// three vertices, two node products, the arithmetic of motion_core.h (shared with the host build the CPU suite checks against the oracle).
// One thread per pixel; neighbouring pixels mostly share a triangle, so the gathers are L1 / L2 hits.  Per covered pixel: 8 B read, 4 B
// written, ~400 fp32 operations and six IEEE divisions (224 of the operations are the two node products, recomputed per pixel here; the
// per-node table the rasteriser uses would remove them) — issue-bound, not yet measured.  Optional: nothing on the cull -> raster ->
// pyramid path depends on it.
#include "kernels.cuh"
#include "motion_core.h"

namespace {

__global__ void __launch_bounds__(256) motion_kernel(const MotionParams p) {
	const size_t n = (size_t)p.W * p.H;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const uint32_t id = (uint32_t)__ldcs(p.vis + i);
		uint32_t packed = 0u;                                     // the attachment's clear value: (0.0, 0.0)
		if (id != VKV_VISBUFFER_CLEAR) {
			const uint32_t drawIndex = id >> VKV_TRIANGLE_BITS, tri = id & ((1u << VKV_TRIANGLE_BITS) - 1u);
			const vkv_MeshletDraw d = p.draws[drawIndex];
			const vkv_Primitive* prim = p.primitives + d.primitiveIndex;
			const vkv_Meshlet* ml = (const vkv_Meshlet*)prim->meshletBuffer + d.meshletIndex;
			const uint8_t* t3 = (const uint8_t*)prim->primitiveIndexBuffer + ml->triangleOffset + tri * 3u;
			const uint32_t* vidx = (const uint32_t*)prim->vertexIndexBuffer + ml->vertexOffset;
			const vkv_Vertex* verts = (const vkv_Vertex*)prim->vertexBuffer;
			const float* T = p.transforms + (size_t)d.transformIndex * 16;
			float mvp[16], prevMvp[16];
			vkv_motion::mul44m(p.camera->viewProjection, T, mvp);
			vkv_motion::mul44m(p.camera->prevViewProjection, T, prevMvp);
			const float* p0 = verts[vidx[t3[0]]].position;
			const float* p1 = verts[vidx[t3[1]]].position;
			const float* p2 = verts[vidx[t3[2]]].position;
			float mv[2];
			vkv_motion::motion_pixel(mvp, prevMvp, p0, p1, p2, (uint32_t)(i % p.W), (uint32_t)(i / p.W), p.W, p.H, mv);
			packed = (uint32_t)vkv_motion::half_rn(mv[0]) | ((uint32_t)vkv_motion::half_rn(mv[1]) << 16);
		}
		p.out[i] = packed;
	}
}

} // namespace

cudaError_t launch_motion(const MotionParams& p, int num_sms, cudaStream_t stream) {
	const size_t n = (size_t)p.W * p.H;
	if (!n) return cudaSuccess;
	size_t grid = (n + 255) / 256;
	if (grid > (size_t)num_sms * 16) grid = (size_t)num_sms * 16;
	motion_kernel<<<(unsigned)grid, 256, 0, stream>>>(p);
	return cudaGetLastError();
}
