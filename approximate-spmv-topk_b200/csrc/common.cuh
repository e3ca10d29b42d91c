// common.cuh -- shared device/host helpers for libtopkspmv (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tks {

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------
// 64-bit ordering keys.  A result is (score, row).  The contract is a total
// order: score descending, then the configured tie-break on the row index
// (evaluation_utils.hpp:52-56 is the "higher index first" variant).
// key = ordered(score) << 32 | (tie_lower ? ~row : row); larger key = better.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float s) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(s);
#else
    union { float f; uint32_t u; } c; c.f = s; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}

__host__ __device__ __forceinline__ uint64_t make_key(uint32_t ordered_score, uint32_t row, int tie_higher) {
    return ((uint64_t)ordered_score << 32) | (uint64_t)(tie_higher ? row : ~row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key, int tie_higher) {
    uint32_t lo = (uint32_t)key;
    return tie_higher ? lo : ~lo;
}
__host__ __device__ __forceinline__ uint32_t key_score(uint64_t key) { return (uint32_t)(key >> 32); }

#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ unsigned lanemask_le() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// 128-bit streaming loads: read-only path, do not allocate in L1 (every matrix
// byte is touched exactly once per query).
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Bitonic sort, DESCENDING, of n (power of two) 64-bit keys in shared memory by
// `nthreads` cooperating threads (tid in [0,nthreads)).  `sync` is __syncwarp()
// for a single warp or __syncthreads() for a whole CTA.
template <typename SyncF>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t *keys, uint32_t n, uint32_t tid, uint32_t nthreads,
                                                  SyncF sync) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            sync();
            for (uint32_t i = tid; i < (n >> 1); i += nthreads) {
                uint32_t lo = 2 * i - (i & (stride - 1));
                uint32_t hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = keys[lo], b = keys[hi];
                bool swap = desc ? (a < b) : (a > b);
                if (swap) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    sync();
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmaticStreamSerialization attribute may
// start while its predecessor in the stream is still running; it must not touch anything the predecessor writes
// before pdl_wait() (which returns once the predecessor grid has completed and its writes are visible).
// pdl_trigger() in the predecessor allows that early start.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#endif  // __CUDACC__

}  // namespace tks
