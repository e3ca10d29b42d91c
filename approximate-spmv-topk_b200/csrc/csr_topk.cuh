// csr_topk.cuh -- fused fp32 CSR Top-K SpMV for sm_100a.
//
// Replaces the reference GPU path  cusparseSpMV / light_spmv  ->  full product
// vector in HBM  ->  thrust::sort_by_key over all N rows  ->  get_topk
// (src/gpu/host_spmv_topk_csr_gpu.cu:171-231, src/gpu/light_spmv.cuh:18-83)
// with ONE streaming pass over the non-zeros whose only HBM output is a short
// list of candidates.  See DESIGN.md for the layout, the roofline and the
// argument that the result equals the exact top-k under the stated total order.
//
// Device layout (built once by tks_upload_csr, see csr_build.cuh):
//   val  [nnz]  fp32, as uploaded
//   colf [nnz]  u32 : bits 0..13 column, bits 14..31 "row delta" = how many rows
//                     the row counter advances AT this element (0 = same row as
//                     the previous non-zero; >= 1 on the first non-zero of a row;
//                     > 1 skips empty rows).  Same 4 bytes as a CSR column index,
//                     so the kernel never reads row_ptr.
//   chunk_start[c], chunk_rb[c] : work units of ~chunk_nnz non-zeros, aligned to
//                     row starts; rb = last non-empty row before the chunk.
//
// Kernels per query:  csr_sample_kernel (threshold from a 0.3 % sample)
//                  -> csr_topk_main_kernel (the HBM stream, > 95 % of the time)
//                  -> select_topk_kernel   (k best of the surviving candidates)
#pragma once

#include "common.cuh"

namespace tks {

constexpr uint32_t kColBits = 14;
constexpr uint32_t kColMask = (1u << kColBits) - 1u;
constexpr uint32_t kMaxDelta = (1u << (32 - kColBits)) - 1u;
constexpr uint32_t kElemsPerLane = 4;                       // one 128-bit load per array
constexpr uint32_t kElemsPerIter = kWarp * kElemsPerLane;   // 128 non-zeros per warp iteration
constexpr uint32_t kSampleIters = 4;                        // sample = first 512 nnz of a chunk
constexpr uint32_t kMainThreads = 512;
constexpr uint32_t kSampleThreads = 256;
constexpr uint32_t kFull = 0xFFFFFFFFu;

struct CsrDevice {
    const float *val;
    const uint32_t *colf;
    const uint64_t *chunk_start;   // n_chunks + 1 entries
    const uint32_t *chunk_rb;      // n_chunks entries
    uint32_t n_chunks;
    uint32_t cols;
    uint32_t row_offset;           // added to every reported row id
};

// Per-query scratch that lives in HBM; zeroed at creation and by the select kernel.
struct RunState {
    uint32_t chunk_counter;   // dynamic scheduler of the main kernel
    uint32_t pool_count;      // candidates appended to the global pool
    uint32_t tau_key;         // ordered-float lower bound on the k-th best score (0 = none)
    uint32_t sample_ticket;   // last-block election of the sample kernel
    uint32_t result_count;
    uint32_t pad[3];
};

// --------------------------------------------------------------------------
// One warp iteration: 128 consecutive non-zeros, 4 per lane.
// Produces, for the lane's FIRST row boundary, the total of the row that ends
// there (T), plus the lane-local pieces needed for rows that start and end
// inside the lane.  All additions are explicit __fadd_rn/__fmul_rn so that the
// sample kernel and the main kernel produce bit-identical row sums.
// --------------------------------------------------------------------------
struct IterState {
    float seg[4];      // inclusive segmented sums inside the lane
    uint32_t d[4];     // row deltas
    float T;           // completed-row total at the lane's first boundary
    uint32_t nf;       // boundaries in this lane
    uint32_t dsum;     // sum of deltas in this lane
    unsigned fm;       // ballot: lanes with >= 1 boundary
};

template <bool MASKED>
__device__ __forceinline__ void csr_iter(const uint4 vraw, const uint4 craw, const float *__restrict__ xs,
                                         uint64_t ebase, uint64_t s, uint64_t e, float carry_in,
                                         float &carry_out, IterState &o) {
    const unsigned lane = lane_id();
    float v[4] = {__uint_as_float(vraw.x), __uint_as_float(vraw.y), __uint_as_float(vraw.z),
                  __uint_as_float(vraw.w)};
    uint32_t c[4] = {craw.x, craw.y, craw.z, craw.w};
    float p[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        o.d[j] = c[j] >> kColBits;
        p[j] = __fmul_rn(v[j], xs[c[j] & kColMask]);
        if (MASKED) {
            bool in = (ebase + j >= s) && (ebase + j < e);
            p[j] = in ? p[j] : 0.0f;
            o.d[j] = in ? o.d[j] : 0u;
        }
    }
    const bool f0 = o.d[0] != 0, f1 = o.d[1] != 0, f2 = o.d[2] != 0, f3 = o.d[3] != 0;
    o.seg[0] = p[0];
    o.seg[1] = f1 ? p[1] : __fadd_rn(o.seg[0], p[1]);
    o.seg[2] = f2 ? p[2] : __fadd_rn(o.seg[1], p[2]);
    o.seg[3] = f3 ? p[3] : __fadd_rn(o.seg[2], p[3]);
    const float head = f0 ? 0.0f : (f1 ? o.seg[0] : (f2 ? o.seg[1] : (f3 ? o.seg[2] : o.seg[3])));
    o.nf = (uint32_t)f0 + (uint32_t)f1 + (uint32_t)f2 + (uint32_t)f3;
    o.dsum = o.d[0] + o.d[1] + o.d[2] + o.d[3];
    o.fm = __ballot_sync(kFull, o.nf != 0);

    // inclusive segmented scan of the lane tails; segments restart at lanes with a boundary
    const unsigned le = o.fm & lanemask_le();
    const int seg_start = le ? (31 - __clz(le)) : 0;
    float I = o.seg[3];
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
        float t = __shfl_up_sync(kFull, I, dlt);
        if ((int)lane - dlt >= seg_start) I = __fadd_rn(I, t);
    }
    float E = __shfl_up_sync(kFull, I, 1);
    if (lane == 0) E = 0.0f;
    if ((o.fm & lanemask_lt()) == 0) E = __fadd_rn(carry_in, E);
    o.T = __fadd_rn(E, head);
    const float I31 = __shfl_sync(kFull, I, 31);
    carry_out = (o.fm == 0) ? __fadd_rn(carry_in, I31) : I31;
}

// Sinks receive completed rows.  emit() is warp-collective.
struct MaxSink {
    float best;
    bool any;
    float tau;   // always -inf: the sample keeps every completed row
    __device__ __forceinline__ void emit(bool pred, float score, uint32_t) {
        if (pred) { best = any ? fmaxf(best, score) : score; any = true; }
    }
};

template <int CAP>
struct PoolSink {
    uint64_t *buf;        // this warp's CAP keys in shared memory
    uint32_t cnt;         // warp-uniform
    uint32_t k;
    float tau;            // current lower bound used by the filter
    uint32_t *tau_key_g;  // global lower bound (atomicMax)
    uint32_t row_offset;
    int tie_higher;

    __device__ __forceinline__ void compact() {
        const unsigned lane = lane_id();
        for (uint32_t i = cnt + lane; i < CAP; i += kWarp) buf[i] = 0ull;
        bitonic_sort_desc(buf, CAP, lane, kWarp, [] { __syncwarp(); });
        if (cnt > k) cnt = k;
        if (cnt == k) {
            const uint32_t tk = key_score(buf[k - 1]);
            uint32_t old = 0;
            if (lane == 0) old = atomicMax(tau_key_g, tk);
            old = __shfl_sync(kFull, old, 0);
            const uint32_t best = old > tk ? old : tk;
            tau = ordered_to_f32(best);
        }
    }
    __device__ __forceinline__ void emit(bool pred, float score, uint32_t row) {
        const unsigned m = __ballot_sync(kFull, pred);
        if (m == 0) return;
        const uint32_t n = __popc(m);
        if (cnt + n > CAP) compact();
        if (pred) buf[cnt + __popc(m & lanemask_lt())] = make_key(f32_to_ordered(score), row + row_offset, tie_higher);
        cnt += n;
        __syncwarp();
    }
};

// Stream one chunk (or its first max_iters iterations) through `sink`.
template <typename Sink>
__device__ __forceinline__ void csr_process_chunk(const CsrDevice &m, const float *__restrict__ xs, uint32_t c,
                                                  uint32_t max_iters, bool flush_tail, Sink &sink) {
    const unsigned lane = lane_id();
    const uint64_t s = m.chunk_start[c], e = m.chunk_start[c + 1];
    if (s >= e) return;
    const uint64_t a0 = s & ~3ull;
    uint64_t n_iter64 = (e - a0 + kElemsPerIter - 1) / kElemsPerIter;
    const bool truncated = n_iter64 > max_iters;
    const uint32_t n_iter = truncated ? max_iters : (uint32_t)n_iter64;
    const uint4 *vp = reinterpret_cast<const uint4 *>(m.val + a0) + lane;
    const uint4 *cp = reinterpret_cast<const uint4 *>(m.colf + a0) + lane;

    uint32_t R = m.chunk_rb[c];   // row in progress (the bogus one before the chunk at first)
    bool first_pending = true;
    float carry = 0.0f;

    // two-deep software prefetch: loads of iterations it+1 and it+2 are in flight while it is reduced
    uint4 v1 = ldg_stream_u4(vp), c1 = ldg_stream_u4(cp);
    uint4 v2 = v1, c2 = c1;
    if (n_iter > 1) { v2 = ldg_stream_u4(vp + kWarp); c2 = ldg_stream_u4(cp + kWarp); }

    for (uint32_t it = 0; it < n_iter; it++) {
        const uint4 cv = v1, cc = c1;
        v1 = v2; c1 = c2;
        if (it + 2 < n_iter) {
            v2 = ldg_stream_u4(vp + (size_t)(it + 2) * kWarp);
            c2 = ldg_stream_u4(cp + (size_t)(it + 2) * kWarp);
        }
        const uint64_t ebase = a0 + (uint64_t)it * kElemsPerIter + lane * kElemsPerLane;
        IterState o;
        float carry_out;
        if (it == 0 || it + 1 == (uint32_t)n_iter64)
            csr_iter<true>(cv, cc, xs, ebase, s, e, carry, carry_out, o);
        else
            csr_iter<false>(cv, cc, xs, ebase, s, e, carry, carry_out, o);
        carry = carry_out;

        const bool pass = (o.nf != 0) && (o.T >= sink.tau);
        const unsigned pm = __ballot_sync(kFull, pass);
        const unsigned im = __ballot_sync(kFull, o.nf >= 2);
        const uint32_t Rtot = __reduce_add_sync(kFull, o.dsum);
        if (pm | im) {
            // row id in progress when entering this lane = R + exclusive prefix of deltas
            uint32_t pre = o.dsum;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                uint32_t t = __shfl_up_sync(kFull, pre, dlt);
                if ((int)lane >= dlt) pre += t;
            }
            const uint32_t Rl = R + (pre - o.dsum);
            const bool bogus = first_pending && (lane == (unsigned)(__ffs(o.fm) - 1));
            sink.emit(pass && !bogus, o.T, Rl);
            if (im) {
                // rows that start AND end inside one lane (length <= 3)
                uint32_t Rj = Rl;
                bool seen = false;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool fj = o.d[j] != 0;
                    const float sv = o.seg[j > 0 ? j - 1 : 0];
                    sink.emit(fj && seen && (sv >= sink.tau), sv, Rj);
                    if (fj) { seen = true; Rj += o.d[j]; }
                }
            }
        }
        if (o.fm) first_pending = false;
        R += Rtot;
    }
    if (flush_tail && !truncated) {
        // the row in progress at the end of the chunk is complete (chunks end on row boundaries)
        sink.emit(lane == 0 && !first_pending && (carry >= sink.tau), carry, R);
    }
}

__device__ __forceinline__ float tau_from_key(uint32_t key) {
    return key == 0 ? -__int_as_float(0x7f800000) : ordered_to_f32(key);
}

// --------------------------------------------------------------------------
// Kernel 1: threshold from a sample.  Warp w reduces the first kSampleIters
// iterations of chunk w*stride exactly as the main kernel will and keeps the
// best completed row.  The k-th largest of these warp maxima is the score of k
// distinct real rows, hence a valid lower bound on the k-th best score.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(kSampleThreads) csr_sample_kernel(CsrDevice m, const float *__restrict__ x,
                                                                     RunState *st, uint32_t *sample_keys,
                                                                     uint32_t n_sample, uint32_t stride,
                                                                     uint32_t k) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float *xs = reinterpret_cast<float *>(smem_raw);
    for (uint32_t i = threadIdx.x; i < m.cols; i += blockDim.x) xs[i] = x[i];
    __syncthreads();
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    if (gw < n_sample) {
        MaxSink sink{0.0f, false, tau_from_key(0)};
        const uint32_t c = gw * stride;
        if (c < m.n_chunks) csr_process_chunk(m, xs, c, kSampleIters, true, sink);
        // warp max
        uint32_t key = sink.any ? f32_to_ordered(sink.best) : 0u;
        key = __reduce_max_sync(kFull, key);
        if (lane_id() == 0) sample_keys[gw] = key;
    }
    // last block picks the k-th largest
    __shared__ uint32_t s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&st->sample_ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);   // reuse (x no longer needed)
    uint32_t n2 = 1;
    while (n2 < n_sample) n2 <<= 1;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x)
        keys[i] = (i < n_sample) ? (uint64_t)ld_relaxed_u32(sample_keys + i) : 0ull;
    bitonic_sort_desc(keys, n2, threadIdx.x, blockDim.x, [] { __syncthreads(); });
    if (threadIdx.x == 0) {
        if (k <= n_sample && keys[k - 1] != 0ull) atomicMax(&st->tau_key, (uint32_t)keys[k - 1]);
        st->sample_ticket = 0;
    }
}

// --------------------------------------------------------------------------
// Kernel 2: the stream.  Persistent CTAs; every warp pulls chunks from a global
// counter, reduces them, and keeps rows with score >= tau in a private
// shared-memory buffer (sorted and cut to k only if it ever fills).
// --------------------------------------------------------------------------
template <int CAP>
__global__ void __launch_bounds__(kMainThreads, 2)
csr_topk_main_kernel(CsrDevice m, const float *__restrict__ x, RunState *st, uint64_t *pool, uint32_t k,
                     int tie_higher) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float *xs = reinterpret_cast<float *>(smem_raw);
    uint64_t *bufs = reinterpret_cast<uint64_t *>(smem_raw + ((m.cols * 4u + 15u) & ~15u));
    for (uint32_t i = threadIdx.x; i < m.cols; i += blockDim.x) xs[i] = x[i];
    __syncthreads();

    const unsigned lane = lane_id();
    PoolSink<CAP> sink;
    sink.buf = bufs + (threadIdx.x / kWarp) * CAP;
    sink.cnt = 0;
    sink.tau = tau_from_key(0);
    sink.k = k;
    sink.tau_key_g = &st->tau_key;
    sink.row_offset = m.row_offset;
    sink.tie_higher = tie_higher;

    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(&st->chunk_counter, 1u);
        c = __shfl_sync(kFull, c, 0);
        if (c >= m.n_chunks) break;
        sink.tau = fmaxf(sink.tau, tau_from_key(ld_relaxed_u32(&st->tau_key)));
        csr_process_chunk(m, xs, c, 0xFFFFFFFFu, true, sink);
    }

    // hand the survivors to the global pool (filtered by the freshest bound)
    const uint32_t tk = ld_relaxed_u32(&st->tau_key);
    __syncwarp();
    for (uint32_t base = 0; base < sink.cnt; base += kWarp) {
        const uint32_t i = base + lane;
        uint64_t key = (i < sink.cnt) ? sink.buf[i] : 0ull;
        const bool keep = (i < sink.cnt) && (key_score(key) >= tk);
        const unsigned mk = __ballot_sync(kFull, keep);
        if (mk) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(&st->pool_count, (uint32_t)__popc(mk));
            pos = __shfl_sync(kFull, pos, 0);
            if (keep) pool[pos + __popc(mk & lanemask_lt())] = key;
        }
    }
}

// --------------------------------------------------------------------------
// Kernel 3: k best keys of a pool, one CTA.  Pools that fit the shared-memory
// sorter are sorted directly; larger ones go through an exact MSB radix select
// first.  Also resets the per-query scratch for the next run.
// --------------------------------------------------------------------------
constexpr uint32_t kSelectThreads = 1024;
constexpr uint32_t kSelectSortCap = 8192;   // 64 KB of keys

__global__ void __launch_bounds__(kSelectThreads)
select_topk_kernel(const uint64_t *__restrict__ pool, const uint32_t *pool_count_ptr, uint32_t pool_count_imm,
                   uint32_t k, int tie_higher, uint64_t *out_keys, uint32_t *out_idx, float *out_val,
                   uint32_t *out_count, RunState *st_reset) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_sel_bin, s_above, s_cnt;
    const uint32_t tid = threadIdx.x;
    const uint32_t n = pool_count_ptr ? *pool_count_ptr : pool_count_imm;

    uint32_t m = 0;   // number of keys staged in shared memory
    if (n <= kSelectSortCap) {
        for (uint32_t i = tid; i < n; i += blockDim.x) keys[i] = pool[i];
        m = n;
    } else {
        // exact radix select of the k-th largest key, 8 bits at a time from the top
        uint64_t prefix = 0, pmask = 0;
        uint32_t need = k < n ? k : n;   // rank still wanted inside the current prefix bucket
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (uint32_t i = tid; i < 256; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < n; i += blockDim.x) {
                uint64_t key = pool[i];
                if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t acc = 0;
                int b = 255;
                for (; b > 0; b--) {
                    if (acc + hist[b] >= need) break;
                    acc += hist[b];
                }
                s_sel_bin = (uint32_t)b;
                s_above = acc;
            }
            __syncthreads();
            prefix |= (uint64_t)s_sel_bin << shift;
            pmask |= 0xFFull << shift;
            need -= s_above;
            __syncthreads();
        }
        // prefix is now the k-th largest key; keys are unique, so exactly min(k,n) keys are >= it
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += blockDim.x) {
            uint64_t key = pool[i];
            if (key >= prefix) {
                uint32_t pos = atomicAdd(&s_cnt, 1u);
                if (pos < kSelectSortCap) keys[pos] = key;
            }
        }
        __syncthreads();
        m = s_cnt < kSelectSortCap ? s_cnt : kSelectSortCap;
    }
    uint32_t n2 = 32;
    while (n2 < m) n2 <<= 1;
    __syncthreads();
    for (uint32_t i = m + tid; i < n2; i += blockDim.x) keys[i] = 0ull;
    bitonic_sort_desc(keys, n2, tid, blockDim.x, [] { __syncthreads(); });
    const uint32_t cnt = m < k ? m : k;
    for (uint32_t i = tid; i < k; i += blockDim.x) {
        const uint64_t key = (i < cnt) ? keys[i] : 0ull;
        out_keys[i] = key;
        out_idx[i] = (i < cnt) ? key_row(key, tie_higher) : 0u;
        out_val[i] = (i < cnt) ? ordered_to_f32(key_score(key)) : 0.0f;
    }
    if (tid == 0) {
        *out_count = cnt;
        if (st_reset) {
            st_reset->result_count = n;   // pool size, for tks_get_stats
            st_reset->chunk_counter = 0;
            st_reset->pool_count = 0;
            st_reset->tau_key = 0;
            st_reset->sample_ticket = 0;
        }
    }
}

}  // namespace tks
