// csr_batched.cuh -- batched multi-query fused fp32 CSR Top-K SpMV for sm_100a (BASELINE config 5).
//
// One pass over the non-zeros serves 32 queries: the matrix is read from HBM once per 32 queries
// instead of once per query.  The reference has no batched mode (its hosts loop `reset(vec)` +
// `operator()` per query, src/gpu/host_spmv_topk_csr_gpu.cu:399-423); this is the north-star's
// "batched multi-query mode amortises one matrix read across queries".
//
// Work decomposition ("lanes are queries"):
//   * the queries of a pass live in shared memory as a table  xT[col][32]  (128 bytes per column, so one
//     non-zero needs ONE conflict-free 128-byte row of it: 32 products per shared-memory wavefront,
//     which is the floor for this problem -- SURVEY 7-H8: the bound is LDS bandwidth, not HBM);
//   * a warp is eight QUADS of 4 lanes; every lane owns 8 queries (two LDS.128 of the table row: queries 4l..4l+3 and
//     16+4l..16+4l+3 for lane l of the quad; odd quads fetch the upper half first, so the two quads of a quarter-warp
//     never meet in a bank) and every quad walks its own chunk of the CSR stream non-zero by non-zero, so a row's
//     score for a query is accumulated sequentially in a single register, in non-zero order, with separate fp32
//     multiply and add -- exactly the arithmetic of the reference gold (gold_algorithms.hpp:203-213), hence
//     bit-identical scores (template FMA=true trades that for fused multiply-adds);
//   * the (value, column) pairs of a quad's stream are fetched with the same coalesced 256-bit loads as
//     the single-query kernel, staged in a 256-byte per-quad shared-memory window and re-read as
//     quad-wide broadcasts (one LDS.128 delivers 4 values or 4 columns to the 4 lanes; a broadcast instruction costs
//     four wavefronts per warp whatever it serves, so with 8 streams per warp instead of 4 -- round 1's octets -- it
//     costs 0.25 wavefronts per non-zero instead of 0.5, on top of the one wavefront of the table row);
//   * a finished row is compared with the query's threshold tau (k-th largest of a sample, as in
//     csr_topk.cuh) and the rare survivors are appended to the query's pool in HBM.
// Kernels per run:  batched_transpose_kernel -> csr_batched_kernel<SAMPLE> -> batched_tau_kernel
//                -> csr_batched_kernel<MAIN> -> select_topk_kernel (one CTA per query).
// If a query's pool overflows (adversarial score order), its count is reported as kPoolOverflow and the
// host re-runs that query through the single-query kernels, which cannot overflow (api.cu).
#pragma once

#include "csr_topk.cuh"

namespace tks {

constexpr uint32_t kBqPerPass = 32;        // queries per pass: 4 lanes x 8 queries
constexpr uint32_t kBThreads = 768;        // 24 warps, one CTA per SM (the table takes most of shared memory)
constexpr uint32_t kBStreams = 8;          // quads (independent streams) per warp
constexpr uint32_t kBStage = 32;           // non-zeros staged per quad per batch (4 lanes x 8)
constexpr uint32_t kBStageBytes = kBStage * 8u;                       // 32 values + 32 column words
constexpr uint32_t kBStageStride = kBStageBytes + 64u;                // neighbouring quads' windows 16 banks apart

struct BatchedArgs {
    const float *xT;          // [npass][cols+1][32]; row `cols` is all zeros (masked elements point there)
    RunState *st;             // [batch]: tau_key, pool_count
    uint64_t *pool;           // [batch][pool_cap]
    uint32_t pool_cap;
    uint32_t batch, npass;
    uint32_t *pass_counter;   // [npass] dynamic chunk scheduler of the main kernel (reset by the select kernel)
    uint32_t *sample_keys;    // [batch][n_sample]
    uint32_t n_sample, stride;
    uint32_t sample_batches;  // staging batches (64 non-zeros each) a sample octet reduces: ~2 % of the matrix in total
    int tie_higher;
};

__host__ __device__ inline size_t batched_table_bytes(uint32_t cols) { return ((size_t)cols + 1u) * kBqPerPass * 4u; }
__host__ __device__ inline size_t batched_smem_bytes(uint32_t cols) {
    return batched_table_bytes(cols) + (size_t)(kBThreads / 4u) * kBStageStride + 128u;   // + slack to align the table to 128 bytes
}

// queries [batch][cols] row-major -> pass tables [npass][cols+1][32], zero padded
__global__ void batched_transpose_kernel(const float *__restrict__ x, uint32_t batch, uint32_t cols, uint32_t npass,
                                         float *__restrict__ xT, int half) {
    const uint32_t n = npass * (cols + 1u) * kBqPerPass;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t q = i % kBqPerPass, r = i / kBqPerPass;
        const uint32_t col = r % (cols + 1u), pass = r / (cols + 1u);
        const uint32_t gq = pass * kBqPerPass + q;
        float v = (col < cols && gq < batch) ? x[(size_t)gq * cols + col] : 0.0f;
        if (half == 1) v = __half2float(__float2half_rn(v));            // 16-bit value modes: the query is rounded like the values
        else if (half == 2) v = __bfloat162float(__float2bfloat16_rn(v));
        xT[i] = v;
    }
}

// k-th largest of every query's sample maxima -> tau_key (one CTA per query)
__global__ void __launch_bounds__(256) batched_tau_kernel(BatchedArgs a, uint32_t k) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *skeys = reinterpret_cast<uint32_t *>(smem_raw);
    __shared__ uint32_t s_bin, s_above[2];
    __shared__ uint32_t hist[(256 / kWarp) * 256];
    const uint32_t q = blockIdx.x;
    for (uint32_t i = threadIdx.x; i < a.n_sample; i += blockDim.x) skeys[i] = a.sample_keys[(size_t)q * a.n_sample + i];
    __syncthreads();
    const uint32_t thr = block_radix_select<uint32_t, false>([&](uint32_t i) { return skeys[i]; }, a.n_sample, k, hist,
                                                             &s_bin, s_above);
    if (threadIdx.x == 0) a.st[q].tau_key = thr;
}

// A finished row reaches at least one of the lane's 8 thresholds (bit q of `mask`): append it to those queries' pools.
// By value on purpose: arrays passed by reference to an out-of-line function would be spilled to local memory.
__device__ __noinline__ void batched_emit(const CsrDevice &m, const BatchedArgs &a, float s0, float s1, float s2, float s3,
                                          float s4, float s5, float s6, float s7, uint32_t mask, uint32_t ord,
                                          uint32_t qlo, uint32_t qhi) {
    const uint32_t row = (m.row_map ? m.row_map[ord] : ord) + m.row_offset;
    while (mask) {
        const int q = __ffs((int)mask) - 1;
        mask &= mask - 1u;
        const float sc = q == 0 ? s0 : q == 1 ? s1 : q == 2 ? s2 : q == 3 ? s3 : q == 4 ? s4 : q == 5 ? s5 : q == 6 ? s6 : s7;
        const uint32_t gq = (q < 4 ? qlo : qhi) + (uint32_t)(q & 3);
        const uint32_t pos = atomicAdd(&a.st[gq].pool_count, 1u);
        if (pos < a.pool_cap) a.pool[(size_t)gq * a.pool_cap + pos] = make_key(f32_to_ordered(sc), row, a.tie_higher);
    }
}

template <bool SAMPLE, bool FMA>
struct BatchedLane {
    float acc[8];      // the lane's 8 queries of the pass: acc[0..3] <- its first table quarter, acc[4..7] <- its second
    float tau[8];      // MAIN: thresholds of this lane's 8 queries (+inf for padding queries)
    float best[8];     // SAMPLE: best completed row so far
    uint32_t ord;      // ordinal of the row in progress
    uint32_t off_first;   // byte offset of the lane's first table quarter (the second one is off_first ^ 64): even quads
                          // own queries 4l..4l+3 then 16+4l..16+4l+3, odd quads the other way round
    bool have_row;

    __device__ __forceinline__ void finish_row(const CsrDevice &m, const BatchedArgs &a, uint32_t qlo, uint32_t qhi) {
        if (have_row) {
            if (SAMPLE) {
#pragma unroll
                for (int q = 0; q < 8; q++) best[q] = fmaxf(best[q], acc[q]);
            } else {
                bool any = false;
#pragma unroll
                for (int q = 0; q < 8; q++) any |= (acc[q] >= tau[q]);
                if (any) {
                    // rare (a few thousand rows of 10^7 per query): kept out of line, the streaming loop stays small
                    uint32_t mask = 0;
#pragma unroll
                    for (int q = 0; q < 8; q++) mask |= (acc[q] >= tau[q]) ? (1u << q) : 0u;
                    batched_emit(m, a, acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7], mask, ord, qlo, qhi);
                }
            }
        }
    }
    __device__ __forceinline__ void madd4(float (&d)[8], int o, float v, const float4 &x) {
        if (FMA) {
            // packed fp32x2 fused multiply-add (FFMA2 on sm_100): two instructions for four queries
            asm("{ .reg .b64 a, b, d;\n\t"
                "mov.b64 a, {%2, %2}; mov.b64 b, {%3, %4}; mov.b64 d, {%0, %1};\n\t"
                "fma.rn.f32x2 d, a, b, d; mov.b64 {%0, %1}, d; }"
                : "+f"(d[o]), "+f"(d[o + 1]) : "f"(v), "f"(x.x), "f"(x.y));
            asm("{ .reg .b64 a, b, d;\n\t"
                "mov.b64 a, {%2, %2}; mov.b64 b, {%3, %4}; mov.b64 d, {%0, %1};\n\t"
                "fma.rn.f32x2 d, a, b, d; mov.b64 {%0, %1}, d; }"
                : "+f"(d[o + 2]), "+f"(d[o + 3]) : "f"(v), "f"(x.z), "f"(x.w));
        } else {
            d[o] = __fadd_rn(d[o], __fmul_rn(v, x.x)); d[o + 1] = __fadd_rn(d[o + 1], __fmul_rn(v, x.y));
            d[o + 2] = __fadd_rn(d[o + 2], __fmul_rn(v, x.z)); d[o + 3] = __fadd_rn(d[o + 3], __fmul_rn(v, x.w));
        }
    }
    __device__ __forceinline__ void step(uint32_t c, float v, uint32_t tab, const CsrDevice &m,
                                         const BatchedArgs &a, uint32_t qlo, uint32_t qhi) {
        // A row start costs ~30 instructions (threshold tests, clearing the sums) against ~25 for the non-zero itself, and
        // left to itself the compiler predicates them into every step.  One vote keeps them out of the two steps in three
        // in which none of the warp's eight streams starts a row.
        const bool start = (int32_t)c < 0;
        if (__any_sync(kFull, start)) {
            if (start) {   // this non-zero starts a row: the row in progress is complete
                finish_row(m, a, qlo, qhi);
#pragma unroll
                for (int q = 0; q < 8; q++) acc[q] = 0.0f;
                ord++;
                have_row = true;
            }
        }
        // 32-bit shared-memory addresses: rows are 128 bytes and 128-byte aligned, so the second quarter is one XOR away
        const uint32_t a1 = tab + (c & 0x7FFFFFFFu) + off_first;
        float4 x1, x2;
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x1.x), "=f"(x1.y), "=f"(x1.z), "=f"(x1.w) : "r"(a1));
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x2.x), "=f"(x2.y), "=f"(x2.z), "=f"(x2.w) : "r"(a1 ^ 64u));
        madd4(acc, 0, v, x1);
        madd4(acc, 4, v, x2);
    }
};

// One quad streams chunk c (or nothing when c >= n_chunks); all eight quads of the warp run the same number
// of batches (the longest of the eight), the shorter ones on neutral elements.
template <bool SAMPLE, bool FMA>
__device__ __forceinline__ void batched_stream(const CsrDevice &m, const BatchedArgs &a, uint32_t tab,
                                               uint8_t *stage, uint32_t c, BatchedLane<SAMPLE, FMA> &L, uint32_t qlo,
                                               uint32_t qhi) {
    const unsigned l4 = lane_id() & 3u;
    uint64_t s = 0, e = 0, a0 = 0;
    uint32_t nb = 0;
    if (c < m.n_chunks) {
        s = m.chunk_start[c];
        e = m.chunk_start[c + 1];
        // a quad's batch is 128 bytes of values: start it on a 128-byte line (TKS experiment r02aa: with 32-byte alignment the
        // L1 asked the L2 for 1.85x the sectors the loads touched)
        a0 = s & ~(uint64_t)(kBStage - 1u);
        nb = (e > s) ? (uint32_t)((e - a0 + kBStage - 1) / kBStage) : 0u;
        L.ord = m.chunk_ord[c] - 1u;
    }
    bool truncated = false;
    if (SAMPLE && nb > a.sample_batches) { nb = a.sample_batches; truncated = true; }
    const uint32_t nb_w = __reduce_max_sync(kFull, nb);
#pragma unroll
    for (int q = 0; q < 8; q++) L.acc[q] = 0.0f;
    L.have_row = false;
    const bool half = m.val_type != 0;   // warp-uniform: 16-bit values (half or bfloat16), widened when staged
    const bool bf16 = m.val_type == 2;
    const uint32_t vshift = half ? 1u : 2u;
    const uint8_t *vp = reinterpret_cast<const uint8_t *>(m.val) + (a0 << vshift) + ((l4 * 8u) << vshift);
    const uint8_t *cp = reinterpret_cast<const uint8_t *>(m.col16 + a0) + l4 * 16u;
    const uint8_t *rp = m.rowbits + (a0 >> 3) + l4;
    const uint32_t zero_off = m.cols * (kBqPerPass * 4u);
    auto load_vals = [&](const uint8_t *p) {
        U32x8 r;
        if (half) {
            const U32x4 h4 = ldg_stream_128(p);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (bf16) {
                    r.w[2 * j] = h4.w[j] << 16;
                    r.w[2 * j + 1] = h4.w[j] & 0xFFFF0000u;
                } else {
                    const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&h4.w[j]));
                    r.w[2 * j] = __float_as_uint(f2.x);
                    r.w[2 * j + 1] = __float_as_uint(f2.y);
                }
            }
        } else {
            r = ldg_stream_256(p);
        }
        return r;
    };
    float *sval = reinterpret_cast<float *>(stage);
    uint32_t *scol = reinterpret_cast<uint32_t *>(stage + kBStage * 4u);

    U32x8 nv;
    U32x4 nc;
    uint32_t nr = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) nv.w[j] = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) nc.w[j] = 0;
    if (nb > 0) { nv = load_vals(vp); nc = ldg_stream_128(cp); nr = ldg_stream_u8(rp); }
    for (uint32_t b = 0; b < nb_w; b++) {
        // column offsets (col * 4) become byte offsets of the table row (col * 128), the row-start bit goes to bit 31
        uint32_t cw[8], vw[8];
        uint32_t lo = 0, hi = 8;
        if (b == 0 || b + 1 >= nb) {
            const int64_t ebase = (int64_t)(a0 + (uint64_t)b * kBStage + l4 * 8u);
            const int64_t l64 = (int64_t)s - ebase, h64 = (int64_t)e - ebase;
            lo = l64 < 0 ? 0u : (l64 > 8 ? 8u : (uint32_t)l64);
            hi = h64 < 0 ? 0u : (h64 > 8 ? 8u : (uint32_t)h64);
            if (b >= nb) { lo = 0; hi = 0; }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t c16 = (j & 1) ? (nc.w[j >> 1] >> 16) : (nc.w[j >> 1] & 0xFFFFu);
            const bool in = ((uint32_t)j >= lo) && ((uint32_t)j < hi);
            cw[j] = in ? ((c16 << 5) | (((nr >> j) & 1u) << 31)) : zero_off;
            vw[j] = in ? nv.w[j] : 0u;
        }
        __syncwarp();   // the previous batch has been consumed
        reinterpret_cast<uint4 *>(sval)[l4 * 2] = make_uint4(vw[0], vw[1], vw[2], vw[3]);
        reinterpret_cast<uint4 *>(sval)[l4 * 2 + 1] = make_uint4(vw[4], vw[5], vw[6], vw[7]);
        reinterpret_cast<uint4 *>(scol)[l4 * 2] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
        reinterpret_cast<uint4 *>(scol)[l4 * 2 + 1] = make_uint4(cw[4], cw[5], cw[6], cw[7]);
        __syncwarp();
        if (b + 1 < nb) {   // next batch in flight while this one is consumed
            nv = load_vals(vp + ((size_t)(b + 1) * kBStage << vshift));
            nc = ldg_stream_128(cp + (size_t)(b + 1) * (kBStage * 2u));
            nr = ldg_stream_u8(rp + (size_t)(b + 1) * (kBStage / 8u));
        }
#pragma unroll 4
        for (uint32_t i = 0; i < kBStage / 4; i++) {
            const float4 v4 = reinterpret_cast<const float4 *>(sval)[i];   // quad-wide broadcast
            const uint4 c4 = reinterpret_cast<const uint4 *>(scol)[i];
            L.step(c4.x, v4.x, tab, m, a, qlo, qhi);
            L.step(c4.y, v4.y, tab, m, a, qlo, qhi);
            L.step(c4.z, v4.z, tab, m, a, qlo, qhi);
            L.step(c4.w, v4.w, tab, m, a, qlo, qhi);
        }
    }
    // chunks end on row boundaries: the row in progress is complete unless the sample cut the chunk short
    if (!truncated) L.finish_row(m, a, qlo, qhi);
    L.have_row = false;
}

// Dynamic shared memory: batched_smem_bytes(cols).
template <bool SAMPLE, bool FMA>
__global__ void __launch_bounds__(kBThreads, 1) csr_batched_kernel(CsrDevice m, BatchedArgs a) {
    extern __shared__ __align__(16) uint8_t smem_dyn[];
    // table rows are 128 bytes and must be 128-byte aligned (the second quarter of a row is addressed as first ^ 64)
    uint8_t *smem_raw = smem_dyn + ((128u - ((uint32_t)__cvta_generic_to_shared(smem_dyn) & 127u)) & 127u);
    const size_t tab_bytes = batched_table_bytes(m.cols);
    const unsigned lane = lane_id(), l4 = lane & 3u, quad = lane >> 2;
    const uint32_t warp = threadIdx.x / kWarp, nwarps = blockDim.x / kWarp;
    uint8_t *stage = smem_raw + tab_bytes + (size_t)(warp * kBStreams + quad) * kBStageStride;
    BatchedLane<SAMPLE, FMA> L;
    const bool odd = (quad & 1u) != 0;
    L.off_first = (odd ? 64u : 0u) + l4 * 16u;
    const uint32_t tab = (uint32_t)__cvta_generic_to_shared(smem_raw);   // the table sits at the start, 128-byte aligned

    for (uint32_t pass = 0; pass < a.npass; pass++) {
        __syncthreads();   // every warp is done with the previous pass's table
        {
            const float4 *src = reinterpret_cast<const float4 *>(a.xT + (size_t)pass * (m.cols + 1u) * kBqPerPass);
            float4 *dst = reinterpret_cast<float4 *>(smem_raw);
            const uint32_t n4 = (m.cols + 1u) * (kBqPerPass / 4u);
            for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
        }
        __syncthreads();
        // first quarter -> acc[0..3], second quarter -> acc[4..7]
        const uint32_t qlo = pass * kBqPerPass + l4 * 4u + (odd ? 16u : 0u), qhi = pass * kBqPerPass + l4 * 4u + (odd ? 0u : 16u);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t gq = (q < 4 ? qlo : qhi) + (uint32_t)(q & 3);
            L.tau[q] = (!SAMPLE && gq < a.batch) ? tau_from_key(ld_relaxed_u32(&a.st[gq].tau_key))
                                                 : __int_as_float(0x7f800000);
            L.best[q] = neg_inf();
        }
        if (SAMPLE) {
            const uint32_t n_groups = (a.n_sample + kBStreams - 1u) / kBStreams;   // one warp reduces 8 samples (one per quad)
            for (uint32_t g = blockIdx.x * nwarps + warp; g < n_groups; g += gridDim.x * nwarps) {
                const uint32_t sidx = g * kBStreams + quad;
                const uint32_t c = (sidx < a.n_sample) ? sidx * a.stride : 0xFFFFFFFFu;
#pragma unroll
                for (int q = 0; q < 8; q++) L.best[q] = neg_inf();
                batched_stream<SAMPLE, FMA>(m, a, tab, stage, c, L, qlo, qhi);
                if (sidx < a.n_sample) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t gq = (q < 4 ? qlo : qhi) + (uint32_t)(q & 3);
                        if (gq < a.batch)
                            a.sample_keys[(size_t)gq * a.n_sample + sidx] =
                                (L.best[q] == neg_inf()) ? 0u : f32_to_ordered(L.best[q]);
                    }
                }
            }
        } else {
            for (;;) {
                uint32_t c0 = 0;
                if (lane == 0) c0 = atomicAdd(&a.pass_counter[pass], kBStreams);
                c0 = __shfl_sync(kFull, c0, 0);
                if (c0 >= m.n_chunks) break;
                batched_stream<SAMPLE, FMA>(m, a, tab, stage, c0 + quad, L, qlo, qhi);
            }
        }
    }
}

}  // namespace tks
