// selftest.cu — device-side self checks of arithmetic building blocks the parity contract leans on.
// vkv_selftest_division: common.cuh's div3_shared (three quotients sharing one refined reciprocal) and its packed f32x2
// form (refined_rcp2 / div_by2, two divisors per instruction) against the `/` operator (IEEE division, what the oracle's
// C++ computes) on pseudo-random operands covering the whole accepted range; and the packed product-then-sum
// add2(mul2(a, b), c) against __fadd_rn(__fmul_rn(a, b), c), i.e. that nothing contracted it into a fused multiply-add.
#include "kernels.cuh"

namespace {

__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
	uint64_t z = (s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
// sign | exponent in [64, 190] (2^-63 .. 2^63) | 23 random mantissa bits; every 8th value gets an extreme / special mantissa
__device__ __forceinline__ float operand(uint64_t r, bool narrow) {
	const uint32_t sign = (uint32_t)(r >> 63) << 31;
	uint32_t e = narrow ? 120u + (uint32_t)((r >> 40) % 15u) : 64u + (uint32_t)((r >> 40) % 127u);
	uint32_t m = (uint32_t)r & 0x7fffffu;
	switch ((r >> 56) & 15u) {
	case 0: m = 0; break;
	case 1: m = 0x7fffffu; break;
	case 2: m = 1; break;
	case 3: m &= 0x7ff000u; break; // few significant bits: exact / halfway quotients become likely
	default: break;
	}
	return __uint_as_float(sign | (e << 23) | m);
}

__global__ void division_selftest_kernel(uint64_t seed, uint32_t itersPerThread, unsigned long long* mismatches, unsigned long long* tested, f2 nz) {
	uint64_t s = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0xD1342543DE82EF95ull);
	unsigned long long bad = 0, n = 0;
	for (uint32_t it = 0; it < itersPerThread; ++it) {
		const bool narrow = (it & 1) != 0;
		const float x = operand(splitmix(s), narrow), y = operand(splitmix(s), narrow), z = operand(splitmix(s), narrow), w = operand(splitmix(s), narrow);
		if (!(div_in_range(x) && div_in_range(y) && div_in_range(z) && div_in_range(w))) continue;
		float qx, qy, qz;
		div3_shared(x, y, z, w, qx, qy, qz);
		const float rx = x / w, ry = y / w, rz = z / w;
		bad += (__float_as_uint(qx) != __float_as_uint(rx)) + (__float_as_uint(qy) != __float_as_uint(ry)) + (__float_as_uint(qz) != __float_as_uint(rz));
		n += 3;
		// packed: (x, y) / (w, z) as one f32x2 division, then (x*y + z, w*z + x) as packed product-then-sum
		const f2 nw = pk(-w, -z);
		const f2 q = div_by2(pk(x, y), nw, refined_rcp2(nw, pk(1.0f, 1.0f)), nz);
		const float px = x / w, py = y / z;
		bad += (__float_as_uint(lo_of(q)) != __float_as_uint(px)) + (__float_as_uint(hi_of(q)) != __float_as_uint(py));
		const f2 ms = add2(mul2(pk(x, w), pk(y, z), nz), pk(z, x));
		const float s0 = __fadd_rn(__fmul_rn(x, y), z), s1 = __fadd_rn(__fmul_rn(w, z), x);
		bad += (__float_as_uint(lo_of(ms)) != __float_as_uint(s0)) + (__float_as_uint(hi_of(ms)) != __float_as_uint(s1));
		n += 4;
	}
	atomicAdd(mismatches, bad);
	atomicAdd(tested, n);
}

} // namespace

cudaError_t launch_division_selftest(uint64_t seed, uint32_t iters, unsigned long long* counters2, unsigned long long negZero2, int num_sms, cudaStream_t stream) {
	division_selftest_kernel<<<num_sms * 8, 256, 0, stream>>>(seed, iters, counters2, counters2 + 1, negZero2);
	return cudaGetLastError();
}
