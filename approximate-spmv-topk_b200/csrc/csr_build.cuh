// csr_build.cuh -- one-time (per matrix) device-side preparation for the fused CSR kernel:
//   * col16 = column * 4 (u16) and the row-start bitmap   (see csr_topk.cuh)
//   * chunk table aligned to row starts
//   * synthetic matrices generated in HBM with the law of the reference generator
//     (src/resources/python/create_matrices.py:84-104), for sizes no MTX file can hold.
// None of this runs per query.
#pragma once

#include "csr_topk.cuh"

namespace tks {

// error bits written by the build kernels
constexpr uint32_t kErrColRange = 1u;     // column index >= cols
constexpr uint32_t kErrPtrOrder = 4u;     // row_ptr not non-decreasing / out of range

// The reference's float_to_half (host_spmv_topk_csr_gpu.cu:152): round to nearest even.
__global__ void csr_vals_to_half_kernel(const float *__restrict__ val, uint64_t nnz, __half *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) out[i] = __float2half_rn(val[i]);
}

__global__ void csr_vals_to_bf16_kernel(const float *__restrict__ val, uint64_t nnz, __nv_bfloat16 *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) out[i] = __float2bfloat16_rn(val[i]);
}

inline float half_bits_to_float(uint16_t bits) {
    __half_raw r;
    r.x = bits;
    return __half2float(__half(r));
}

template <typename P>
__global__ void csr_copy_cols_kernel(const uint32_t *__restrict__ idx, uint64_t nnz, uint32_t cols,
                                     uint16_t *__restrict__ col16, uint32_t *err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < nnz; i += stride) {
        uint32_t c = idx[i];
        bad |= (c >= cols);
        col16[i] = (uint16_t)((c << 2) & kColOffMask);
    }
    if (bad) atomicOr(err, kErrColRange);
}

// col12: the column offsets (column * 4 <= 4092) of 8 consecutive non-zeros packed into three 32-bit words, element j at
// bits [12j, 12j + 12) -- one thread per group of 8; the tail group is padded with zeros
__global__ void csr_pack_cols12_kernel(const uint16_t *__restrict__ col16, uint64_t nnz, uint32_t *__restrict__ col12) {
    const uint64_t ngroups = (nnz + 7) / 8;
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; g < ngroups; g += stride) {
        uint64_t lo = 0, hi = 0;   // 96 bits
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint64_t i = g * 8 + j;
            const uint64_t c = (i < nnz) ? (uint64_t)(col16[i] & 0xFFFu) : 0ull;
            const int bit = 12 * j;
            if (bit >= 64) hi |= c << (bit - 64);
            else {
                lo |= c << bit;
                if (bit + 12 > 64) hi |= c >> (64 - bit);
            }
        }
        col12[g * 3 + 0] = (uint32_t)lo;
        col12[g * 3 + 1] = (uint32_t)(lo >> 32);
        col12[g * 3 + 2] = (uint32_t)hi;
    }
}

// One thread per row: tag the row's first non-zero; count non-empty rows.
template <typename P>
__global__ void csr_mark_rows_kernel(const P *__restrict__ ptr, uint64_t rows, uint64_t nnz,
                                     uint32_t *__restrict__ rowbits32, uint32_t *nonempty_flag, uint32_t *err) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const uint64_t b = ptr[r], e = ptr[r + 1];
    // row_ptr must run from 0 to nnz: non-zeros in front of ptr[0] or behind ptr[rows] would silently join a
    // neighbouring row's score
    const bool bad_ends = (r == 0 && b != 0) || (r + 1 == rows && e != nnz);
    if (e < b || e > nnz || bad_ends) { atomicOr(err, kErrPtrOrder); nonempty_flag[r] = 0; return; }
    nonempty_flag[r] = (b != e) ? 1u : 0u;
    if (b != e) atomicOr(&rowbits32[b >> 5], 1u << (b & 31u));   // little-endian: bit b%8 of byte b/8
}

// row_map[ordinal] = row for non-empty rows, given ord = exclusive scan of the non-empty flags (as u64).
__global__ void csr_row_map_kernel(const uint32_t *__restrict__ nonempty_flag, const uint64_t *__restrict__ ord,
                                   uint64_t rows, uint32_t *__restrict__ row_map) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows && nonempty_flag[r]) row_map[ord[r]] = (uint32_t)r;
}

// One thread per chunk: first row starting at or after the chunk's nominal offset, and its ordinal among non-empty rows.
// The first n_big chunks are chunk_nnz non-zeros long, the rest (the tail of the stream) chunk_small: every chunk start
// costs a warp three dependent round trips (scheduler atomic, this table, first loads), so the work units are large --
// but the persistent warps should still run dry together, so the last tenth of the stream is handed out in quarters.
template <typename P>
__global__ void csr_chunk_table_kernel(const P *__restrict__ ptr, uint64_t rows, uint64_t nnz, uint32_t chunk_nnz,
                                       uint32_t n_big, uint32_t chunk_small, uint32_t n_chunks,
                                       const uint64_t *__restrict__ ord, uint64_t *__restrict__ chunk_start,
                                       uint32_t *__restrict__ chunk_ord) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    if (c == n_chunks) { chunk_start[c] = nnz; return; }
    const uint64_t target = c < n_big ? (uint64_t)c * chunk_nnz : (uint64_t)n_big * chunk_nnz + (uint64_t)(c - n_big) * chunk_small;
    uint64_t lo = 0, hi = rows;   // lower_bound over ptr[0..rows)
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if ((uint64_t)ptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    // lo = first row with ptr[row] >= target (or rows); empty rows at lo share its offset
    chunk_start[c] = (lo < rows) ? (uint64_t)ptr[lo] : nnz;
    // ordinal of the first non-empty row >= lo: ord[] is an exclusive count of non-empty rows (nullptr: none empty)
    chunk_ord[c] = ord ? (uint32_t)((lo < rows) ? ord[lo] : ord[rows]) : (uint32_t)lo;
}

// ---------------------------------------------------------------------------
// Synthetic generator (create_matrices.py): degree law, sorted random columns
// with replacement, U[0,1) values divided by the row's L2 norm.  Counter-based
// RNG so every (seed,row) is reproducible independent of launch geometry.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

constexpr uint32_t kGenMaxDegree = 512;

__host__ __device__ __forceinline__ uint32_t synth_degree(uint64_t seed, uint64_t grow, uint32_t avg, int dist) {
    const uint64_t h0 = splitmix64(seed ^ (grow * 0xD1B54A32D192ED03ull));
    uint32_t deg;
    if (dist == 0) {
        // uniform on [avg/2, int(1.5 avg)]   (create_matrices.py:84-86)
        const uint32_t lo = avg / 2, hi = (uint32_t)(avg * 1.5);
        deg = lo + (uint32_t)(splitmix64(h0) % (uint64_t)(hi - lo + 1));
    } else {
        // Gamma(shape 3, scale avg/3) = sum of three exponentials; int() then max(.,1)  (:31,:91)
        const double u1 = 1.0 - u01(splitmix64(h0 + 1)), u2 = 1.0 - u01(splitmix64(h0 + 2)),
                     u3 = 1.0 - u01(splitmix64(h0 + 3));
        const double g = -log(u1 * u2 * u3) * ((double)avg / 3.0);
        deg = (uint32_t)g;
        if (deg < 1) deg = 1;
    }
    return deg > kGenMaxDegree ? kGenMaxDegree : deg;
}

__global__ void synth_degree_kernel(uint64_t rows, uint64_t row_offset, uint64_t seed, uint32_t avg, int dist,
                                    uint32_t *__restrict__ deg) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) deg[r] = synth_degree(seed, r + row_offset, avg, dist);
}

// Exclusive scan of deg[rows] into ptr[rows+1] (u64) in three simple phases.
constexpr uint32_t kScanBlock = 1024;
__global__ void scan_block_sums_kernel(const uint32_t *__restrict__ deg, uint64_t rows, uint64_t *block_sums) {
    __shared__ uint64_t sh[kScanBlock / 32];
    uint64_t r = (uint64_t)blockIdx.x * kScanBlock + threadIdx.x;
    uint64_t v = r < rows ? deg[r] : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint64_t t = sh[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(kFull, t, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
    }
}
__global__ void scan_block_offsets_kernel(uint64_t *block_sums, uint32_t n_blocks) {
    // single thread: n_blocks <= a few hundred thousand, setup only
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint64_t acc = 0;
        for (uint32_t i = 0; i < n_blocks; i++) { uint64_t t = block_sums[i]; block_sums[i] = acc; acc += t; }
        block_sums[n_blocks] = acc;
    }
}
__global__ void scan_finish_kernel(const uint32_t *__restrict__ deg, uint64_t rows, const uint64_t *block_sums,
                                   uint64_t *__restrict__ ptr) {
    __shared__ uint64_t sh[kScanBlock];
    uint64_t r = (uint64_t)blockIdx.x * kScanBlock + threadIdx.x;
    sh[threadIdx.x] = r < rows ? deg[r] : 0;
    __syncthreads();
    for (uint32_t o = 1; o < kScanBlock; o <<= 1) {
        uint64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (r < rows) ptr[r] = block_sums[blockIdx.x] + sh[threadIdx.x] - deg[r];
    if (r == rows - 1) ptr[rows] = block_sums[blockIdx.x] + sh[threadIdx.x];
}

__global__ void synth_fill_kernel(uint64_t rows, uint64_t row_offset, uint64_t seed, uint32_t cols,
                                  const uint64_t *__restrict__ ptr, uint32_t *__restrict__ idx,
                                  float *__restrict__ val) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const uint64_t b = ptr[r];
    const uint32_t deg = (uint32_t)(ptr[r + 1] - b);
    const uint64_t hr = splitmix64((seed + 0x5851F42D4C957F2Dull) ^ ((r + row_offset) * 0x2545F4914F6CDD1Dull));
    uint16_t cbuf[kGenMaxDegree];
    double nrm = 0.0;
    for (uint32_t j = 0; j < deg; j++) {
        const uint64_t h = splitmix64(hr + 2 * j);
        uint16_t cj = (uint16_t)(h % cols);
        // insertion sort (sorted(np.random.randint(...)), create_matrices.py:45)
        int q = (int)j - 1;
        while (q >= 0 && cbuf[q] > cj) { cbuf[q + 1] = cbuf[q]; q--; }
        cbuf[q + 1] = cj;
        const double v = u01(splitmix64(hr + 2 * j + 1));
        nrm += v * v;
    }
    const double inv = nrm > 0.0 ? 1.0 / sqrt(nrm) : 0.0;
    for (uint32_t j = 0; j < deg; j++) {
        idx[b + j] = cbuf[j];
        val[b + j] = (float)(u01(splitmix64(hr + 2 * j + 1)) * inv);
    }
}

}  // namespace tks
