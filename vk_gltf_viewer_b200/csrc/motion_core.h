// motion_core.h — the per-pixel arithmetic of the motion-vector pass, shared by the CUDA kernel (motion.cu) and by a host build of the same
// source (tests/motion_core_shim.cpp: g++ -ffp-contract=off) that the CPU suite compares bit for bit with the oracle's independent
// restatement (orc_motion_vectors).  nvcc compiles this file with -fmad=false and IEEE division (-prec-div=true), g++ with
// -ffp-contract=off: every fp32 operation below is individually rounded on both sides, so host and device produce the same bits.
//
// What it computes: the second colour attachment of the reference's visbuffer pass (R16G16_SFLOAT "Motion vectors", application.cpp:
// 250-267, cleared to 0 at :786-799) — visbuffer.frag.glsl:38
//     motionVectors = ((prevPosition.xy / prevPosition.w) * 0.5f - 0.5f) - ((position.xy / position.w) * 0.5f - 0.5f);
// with position = viewProjection * transform * vertex and prevPosition = prevViewProjection * transform * vertex (visbuffer.mesh.glsl:
// 44-45,61-63) interpolated perspective-correctly over the triangle.  The reference evaluates it per fragment while rasterising; here it is
// a pass over the finished visbuffer (one evaluation per PIXEL instead of per fragment): the pixel's id names the triangle, the three
// vertices are fetched again and the varyings are interpolated at the pixel centre.
//
// Interpolation (Vulkan: perspective-correct barycentrics; their arithmetic is implementation-defined): homogeneous edge functions in
// (x, y, w) clip space — lambda_i(n) = cofactor_i . (n.x, n.y, 1) for the pixel centre n in NDC, the rows of adj([x y w]) — which need no
// clipping for vertices behind the eye.  A varying interpolates as sum(lambda_i f_i) / sum(lambda_i); the shader only uses ratios of two
// components of the same varying, so the common divisor is dropped.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VKV_HD __host__ __device__ __forceinline__
#else
#define VKV_HD inline
#endif

namespace vkv_motion {

struct V4 { float x, y, z, w; };

// mat4 (column-major m[c*4+r]) * vec4 : ((c0*x + c1*y) + c2*z) + c3*w — the path's arithmetic policy
VKV_HD V4 mul44(const float* m, float x, float y, float z, float w) {
	V4 r;
	r.x = ((m[0] * x + m[4] * y) + m[8] * z) + m[12] * w;
	r.y = ((m[1] * x + m[5] * y) + m[9] * z) + m[13] * w;
	r.z = ((m[2] * x + m[6] * y) + m[10] * z) + m[14] * w;
	r.w = ((m[3] * x + m[7] * y) + m[11] * z) + m[15] * w;
	return r;
}
VKV_HD void mul44m(const float* a, const float* b, float* out) { // out = a * b, column by column (visbuffer.mesh.glsl:44-45)
	for (int c = 0; c < 4; ++c) {
		const V4 col = mul44(a, b[c * 4 + 0], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]);
		out[c * 4 + 0] = col.x; out[c * 4 + 1] = col.y; out[c * 4 + 2] = col.z; out[c * 4 + 3] = col.w;
	}
}

// fp32 -> fp16, round to nearest even (the R16G16_SFLOAT store), integer arithmetic only so that every compiler agrees; NaN -> 0x7FFF
VKV_HD uint16_t half_rn(float f) {
	union { float f; uint32_t u; } v;
	v.f = f;
	const uint32_t sign = (v.u >> 16) & 0x8000u, a = v.u & 0x7fffffffu;
	if (a > 0x7f800000u) return 0x7fffu;                         // NaN
	if (a >= 0x47800000u) return (uint16_t)(sign | 0x7c00u);     // >= 65536 (or inf): beyond the largest half even before rounding -> inf
	if (a < 0x33000001u) return (uint16_t)sign;                  // <= 2^-25: rounds to zero (2^-25 itself is a tie to even = 0)
	uint32_t mant = a & 0x007fffffu, shift;
	int32_t e = (int32_t)(a >> 23) - 127;                        // unbiased exponent
	uint32_t h;
	if (e < -14) { // subnormal half: value = mant24 * 2^(e-23), unit 2^-24
		mant |= 0x00800000u;
		shift = (uint32_t)(-1 - e);                              // 14 .. 24: bits dropped so that the result counts units of 2^-24
		const uint32_t q = mant >> shift, rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1u);
		h = q + ((rem > halfway || (rem == halfway && (q & 1u))) ? 1u : 0u);
	} else {
		const uint32_t q = ((uint32_t)(e + 15) << 10) | (mant >> 13), rem = mant & 0x1fffu;
		h = q + ((rem > 0x1000u || (rem == 0x1000u && (q & 1u))) ? 1u : 0u); // a carry out of the mantissa bumps the exponent (up to inf): correct
	}
	return (uint16_t)(sign | h);
}

// One pixel.  mvp / prevMvp: the node's two products; p0..p2: the triangle's object-space positions; (px, py): pixel coordinates;
// (W, H): render resolution.  out: the two fp32 components before the fp16 store.
VKV_HD void motion_pixel(const float* mvp, const float* prevMvp, const float* p0, const float* p1, const float* p2, uint32_t px, uint32_t py,
                         uint32_t W, uint32_t H, float out[2]) {
	const V4 a0 = mul44(mvp, p0[0], p0[1], p0[2], 1.0f), a1 = mul44(mvp, p1[0], p1[1], p1[2], 1.0f), a2 = mul44(mvp, p2[0], p2[1], p2[2], 1.0f);
	const V4 b0 = mul44(prevMvp, p0[0], p0[1], p0[2], 1.0f), b1 = mul44(prevMvp, p1[0], p1[1], p1[2], 1.0f), b2 = mul44(prevMvp, p2[0], p2[1], p2[2], 1.0f);
	// pixel centre in NDC: the inverse of the viewport transform sx = ndc * (W/2) + W/2
	const float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
	const float nx = (((float)px + 0.5f) - hw) / hw, ny = (((float)py + 0.5f) - hh) / hh;
	// homogeneous edge functions: lambda_i = cofactor row i of [[x0 x1 x2] [y0 y1 y2] [w0 w1 w2]] applied to (nx, ny, 1)
	const float l0 = ((a1.y * a2.w - a1.w * a2.y) * nx + (a1.w * a2.x - a1.x * a2.w) * ny) + (a1.x * a2.y - a1.y * a2.x);
	const float l1 = ((a2.y * a0.w - a2.w * a0.y) * nx + (a2.w * a0.x - a2.x * a0.w) * ny) + (a2.x * a0.y - a2.y * a0.x);
	const float l2 = ((a0.y * a1.w - a0.w * a1.y) * nx + (a0.w * a1.x - a0.x * a1.w) * ny) + (a0.x * a1.y - a0.y * a1.x);
	// the two varyings, up to the common divisor (l0 + l1 + l2)
	const float Px = (l0 * a0.x + l1 * a1.x) + l2 * a2.x, Py = (l0 * a0.y + l1 * a1.y) + l2 * a2.y, Pw = (l0 * a0.w + l1 * a1.w) + l2 * a2.w;
	const float Qx = (l0 * b0.x + l1 * b1.x) + l2 * b2.x, Qy = (l0 * b0.y + l1 * b1.y) + l2 * b2.y, Qw = (l0 * b0.w + l1 * b1.w) + l2 * b2.w;
	// visbuffer.frag.glsl:38
	out[0] = ((Qx / Qw) * 0.5f - 0.5f) - ((Px / Pw) * 0.5f - 0.5f);
	out[1] = ((Qy / Qw) * 0.5f - 0.5f) - ((Py / Pw) * 0.5f - 0.5f);
}

} // namespace vkv_motion
