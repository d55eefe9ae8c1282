// hiz_tile.cuh — the exact-2x mips of one 64x16-pixel visbuffer tile, in registers (shared by hiz.cu and strips.cu).
// hiz_reduce.comp.glsl:21-31 for the levels whose source is exactly twice the destination in both axes: the sampler footprint is
// the aligned 2x2 quad, so mip k of the tile is a min over 2^(k+1) x 2^(k+1) source pixels — register mins down a lane's two
// columns, shfl.xor across lanes.  Lane l owns source columns 2l, 2l+1 (one 16-byte load per row, 16 loads in flight).
#pragma once
#include "common.cuh"

constexpr int kTileW = 64, kTileH = 16;
constexpr int kTileSlots = 15;          // mip texels one lane can own in a tile: 8 + 4 + 2 + 1 // source pixels per warp tile; yields 32x8, 16x4, 8x2, 4x1 texels of mips 0..3

struct HizTileGeo {
	uint32_t W, H, E;          // render resolution, number of exact levels (1..4)
	uint32_t off[4], w[4];     // pyramid offsets / row strides of mips 0..3
};

__device__ __forceinline__ float hiz_min4(float a, float b, float c, float d) { return gmin(gmin(gmin(a, b), c), d); }

// the tile's 16 rows of this lane's two keys; rows / columns outside the image read as key 0 (replaced by +inf in the reduction)
__device__ __forceinline__ void hiz_tile_load(const unsigned long long* __restrict__ vis, const HizTileGeo& g, uint32_t tx, uint32_t ty, uint32_t lane,
                                              ulonglong2 (&v)[kTileH]) {
	const uint32_t x0 = tx * kTileW + lane * 2, y0 = ty * kTileH;
	const bool colIn = x0 < g.W; // W is even whenever E >= 1, so the pair is in or out together
#pragma unroll
	for (int r = 0; r < kTileH; ++r) {
		v[r] = make_ulonglong2(0ull, 0ull);
		if (colIn && y0 + r < g.H) v[r] = __ldcs((const ulonglong2*)(vis + (size_t)(y0 + r) * g.W + x0));
	}
}

// store(slot, index into the pyramid, value) is called for every texel of mips 0..E-1 this lane owns; slot is a compile-time
// constant after unrolling (mip 0: 0..7, mip 1: 8..11, mip 2: 12..13, mip 3: 14), so a caller may keep per-slot state in registers
template <class Store>
__device__ __forceinline__ void hiz_tile_reduce(const ulonglong2 (&v)[kTileH], const HizTileGeo& g, uint32_t tx, uint32_t ty, uint32_t lane, Store&& store) {
	const uint32_t x0 = tx * kTileW + lane * 2, y0 = ty * kTileH;
	const bool colIn = x0 < g.W;
	float m0[8];
#pragma unroll
	for (int r = 0; r < 8; ++r) {
		const bool in = colIn && (y0 + 2 * r + 1 < g.H);
		const float a = depth_of_key(v[2 * r].x), b = depth_of_key(v[2 * r].y);
		const float c = depth_of_key(v[2 * r + 1].x), d = depth_of_key(v[2 * r + 1].y);
		m0[r] = in ? hiz_min4(a, b, c, d) : __int_as_float(0x7f800000);
	}
	{ // mip 0: 32 x 8 per tile
		const uint32_t mx = tx * 32 + lane, my0 = ty * 8;
#pragma unroll
		for (int r = 0; r < 8; ++r)
			if (mx < (g.W >> 1) && my0 + r < (g.H >> 1)) store(r, g.off[0] + (my0 + r) * g.w[0] + mx, m0[r]);
	}
	if (g.E < 2) return;
	float m1[4];
#pragma unroll
	for (int r = 0; r < 4; ++r) {
		const float t = gmin(m0[2 * r], m0[2 * r + 1]);
		m1[r] = gmin(t, __shfl_xor_sync(0xffffffffu, t, 1));
	}
	if ((lane & 1) == 0) {
		const uint32_t mx = tx * 16 + (lane >> 1), my0 = ty * 4;
#pragma unroll
		for (int r = 0; r < 4; ++r)
			if (mx < (g.W >> 2) && my0 + r < (g.H >> 2)) store(8 + r, g.off[1] + (my0 + r) * g.w[1] + mx, m1[r]);
	}
	if (g.E < 3) return;
	float m2[2];
#pragma unroll
	for (int r = 0; r < 2; ++r) {
		const float t = gmin(m1[2 * r], m1[2 * r + 1]);
		m2[r] = gmin(t, __shfl_xor_sync(0xffffffffu, t, 2));
	}
	if ((lane & 3) == 0) {
		const uint32_t mx = tx * 8 + (lane >> 2), my0 = ty * 2;
#pragma unroll
		for (int r = 0; r < 2; ++r)
			if (mx < (g.W >> 3) && my0 + r < (g.H >> 3)) store(12 + r, g.off[2] + (my0 + r) * g.w[2] + mx, m2[r]);
	}
	if (g.E < 4) return;
	{
		const float t = gmin(m2[0], m2[1]);
		const float m3 = gmin(t, __shfl_xor_sync(0xffffffffu, t, 4));
		if ((lane & 7) == 0) {
			const uint32_t mx = tx * 4 + (lane >> 3), my = ty;
			if (mx < (g.W >> 4) && my < (g.H >> 4)) store(14, g.off[3] + my * g.w[3] + mx, m3);
		}
	}
}

// The pyramid indices of the texels hiz_tile_reduce will hand to `store` for this (tile, lane), without the data: visit(slot, index).
// Mirrors hiz_tile_reduce's conditions exactly (same slots, same bounds).
template <class Visit>
__device__ __forceinline__ void hiz_tile_slots(const HizTileGeo& g, uint32_t tx, uint32_t ty, uint32_t lane, Visit&& visit) {
	{
		const uint32_t mx = tx * 32 + lane, my0 = ty * 8;
#pragma unroll
		for (int r = 0; r < 8; ++r)
			if (mx < (g.W >> 1) && my0 + r < (g.H >> 1)) visit(r, g.off[0] + (my0 + r) * g.w[0] + mx);
	}
	if (g.E >= 2 && (lane & 1) == 0) {
		const uint32_t mx = tx * 16 + (lane >> 1), my0 = ty * 4;
#pragma unroll
		for (int r = 0; r < 4; ++r)
			if (mx < (g.W >> 2) && my0 + r < (g.H >> 2)) visit(8 + r, g.off[1] + (my0 + r) * g.w[1] + mx);
	}
	if (g.E >= 3 && (lane & 3) == 0) {
		const uint32_t mx = tx * 8 + (lane >> 2), my0 = ty * 2;
#pragma unroll
		for (int r = 0; r < 2; ++r)
			if (mx < (g.W >> 3) && my0 + r < (g.H >> 3)) visit(12 + r, g.off[2] + (my0 + r) * g.w[2] + mx);
	}
	if (g.E >= 4 && (lane & 7) == 0) {
		const uint32_t mx = tx * 4 + (lane >> 3), my = ty;
		if (mx < (g.W >> 4) && my < (g.H >> 4)) visit(14, g.off[3] + my * g.w[3] + mx);
	}
}
