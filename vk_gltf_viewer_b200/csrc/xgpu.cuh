// xgpu.cuh — cross-GPU signalling over CUDA-IPC peer memory (merge.cu's barrier, and the barriers fused into strips.cu / hiz.cu).
// Every rank owns kMaxRanks u32 slots; rank r signals rank t by writing the current epoch into slot [r] of t's array, and waits by
// polling its OWN array (local memory).  Epochs only grow, and a peer can be at most one barrier ahead, so ">= epoch" is arrival.
#pragma once
#include "kernels.cuh"

__device__ __forceinline__ unsigned long long xgpu_timer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ void xgpu_st_release(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t xgpu_ld_acquire(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
// called by threads 0..nranks-1 of ONE block per rank: thread t tells rank t that this rank has reached `epoch`
__device__ __forceinline__ void xgpu_signal(uint32_t* const* flags, int rank, int t, uint32_t epoch) {
	__threadfence_system();
	xgpu_st_release(flags[t] + rank, epoch);
}
// called by threads 0..nranks-1 of any block: thread t waits for rank t; false = timed out (reported through *error)
__device__ __forceinline__ bool xgpu_wait(const uint32_t* localFlags, int t, uint32_t epoch, unsigned long long timeout_ns, uint32_t* error) {
	const unsigned long long t0 = xgpu_timer_ns();
	while ((int32_t)(xgpu_ld_acquire(localFlags + t) - epoch) < 0) {
		if (xgpu_timer_ns() - t0 > timeout_ns) { *error = 1u; return false; } // a peer never arrived: report instead of hanging the GPU
		__nanosleep(100);
	}
	return true;
}
