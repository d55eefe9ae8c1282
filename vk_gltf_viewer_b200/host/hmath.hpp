// hmath.hpp — small fp32 vector/matrix helpers for the HOST input generators.
// Operation order follows the libraries the reference's host code uses (glm for the camera,
// fastgltf::math for node transforms) so the uploaded matrices are the same bytes; built with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace vkvh {

struct vec3 {
	float x = 0, y = 0, z = 0;
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            // glm compute_dot<vec3>
inline vec3 cross(vec3 x, vec3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }               // glm: v * inversesqrt(dot(v,v))

struct mat4 { // column-major, m[c*4+r]
	float m[16];
	float* col(int c) { return m + c * 4; }
	const float* col(int c) const { return m + c * 4; }
};
inline mat4 identity() {
	mat4 r{};
	r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
	return r;
}
// glm::operator*(mat4, mat4): Result[c] = ((A0*B[c][0] + A1*B[c][1]) + A2*B[c][2]) + A3*B[c][3]
inline mat4 mul(const mat4& a, const mat4& b) {
	mat4 r;
	for (int c = 0; c < 4; ++c)
		for (int i = 0; i < 4; ++i)
			r.m[c * 4 + i] = ((a.m[0 + i] * b.m[c * 4 + 0] + a.m[4 + i] * b.m[c * 4 + 1]) + a.m[8 + i] * b.m[c * 4 + 2]) + a.m[12 + i] * b.m[c * 4 + 3];
	return r;
}

// glm::perspectiveRH_ZO (glm/ext/matrix_clip_space.inl:233-247)
inline mat4 perspectiveRH_ZO(float fovy, float aspect, float zNear, float zFar) {
	const float tanHalfFovy = std::tan(fovy / 2.0f);
	mat4 r{};
	r.m[0] = 1.0f / (aspect * tanHalfFovy);
	r.m[5] = 1.0f / tanHalfFovy;
	r.m[10] = zFar / (zNear - zFar);
	r.m[11] = -1.0f;
	r.m[14] = -(zFar * zNear) / (zFar - zNear);
	return r;
}
// glm::lookAtRH (glm/ext/matrix_transform.inl:153-173)
inline mat4 lookAtRH(vec3 eye, vec3 center, vec3 up) {
	const vec3 f = normalize(center - eye);
	const vec3 s = normalize(cross(f, up));
	const vec3 u = cross(s, f);
	mat4 r = identity();
	r.m[0] = s.x; r.m[4] = s.y; r.m[8] = s.z;
	r.m[1] = u.x; r.m[5] = u.y; r.m[9] = u.z;
	r.m[2] = -f.x; r.m[6] = -f.y; r.m[10] = -f.z;
	r.m[12] = -dot(s, eye); r.m[13] = -dot(u, eye); r.m[14] = dot(f, eye);
	return r;
}

// fastgltf::math::translate / rotate / scale (fastgltf/math.hpp:816-837) and asMatrix (:587-604)
inline mat4 translate(const mat4& m, const float t[3]) {
	mat4 r = m;
	for (int i = 0; i < 4; ++i) r.m[12 + i] = ((m.m[i] * t[0] + m.m[4 + i] * t[1]) + m.m[8 + i] * t[2]) + m.m[12 + i];
	return r;
}
inline mat4 rotate(const mat4& m, const float q[4]) { // q = x,y,z,w
	const float x = q[0], y = q[1], z = q[2], w = q[3];
	mat4 R = identity();
	R.m[0] = 1.0f - 2.0f * (y * y + z * z); R.m[1] = 2.0f * (x * y + w * z); R.m[2] = 2.0f * (x * z - w * y);
	R.m[4] = 2.0f * (x * y - w * z); R.m[5] = 1.0f - 2.0f * (x * x + z * z); R.m[6] = 2.0f * (y * z + w * x);
	R.m[8] = 2.0f * (x * z + w * y); R.m[9] = 2.0f * (y * z - w * x); R.m[10] = 1.0f - 2.0f * (x * x + y * y);
	// fastgltf mat*mat: column c = sum_k m.col(k) * R[c][k], left to right
	mat4 r;
	for (int c = 0; c < 4; ++c)
		for (int i = 0; i < 4; ++i)
			r.m[c * 4 + i] = ((m.m[i] * R.m[c * 4] + m.m[4 + i] * R.m[c * 4 + 1]) + m.m[8 + i] * R.m[c * 4 + 2]) + m.m[12 + i] * R.m[c * 4 + 3];
	return r;
}
inline mat4 scale(const mat4& m, const float s[3]) {
	mat4 r = m;
	for (int i = 0; i < 4; ++i) { r.m[i] = m.m[i] * s[0]; r.m[4 + i] = m.m[4 + i] * s[1]; r.m[8 + i] = m.m[8 + i] * s[2]; }
	return r;
}

// splitmix64 (deterministic PRNG for the synthetic scenes)
struct SplitMix64 {
	uint64_t s;
	explicit SplitMix64(uint64_t seed) : s(seed) {}
	uint64_t next() {
		uint64_t z = (s += 0x9E3779B97F4A7C15ull);
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		return z ^ (z >> 31);
	}
	float unit() { return (float)(next() >> 40) * (1.0f / 16777216.0f); } // [0,1)
	float range(float lo, float hi) { return lo + (hi - lo) * unit(); }
};

} // namespace vkvh
