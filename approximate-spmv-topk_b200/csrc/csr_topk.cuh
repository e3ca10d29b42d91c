// csr_topk.cuh -- fused fp32 CSR Top-K SpMV for sm_100a.
//
// Replaces the reference GPU path  cusparseSpMV / light_spmv  ->  full product
// vector in HBM  ->  thrust::sort_by_key over all N rows  ->  get_topk
// (src/gpu/host_spmv_topk_csr_gpu.cu:171-231, src/gpu/light_spmv.cuh:18-83)
// with ONE streaming pass over the non-zeros whose only HBM output is a short
// list of candidates.  See DESIGN.md for the layout, the roofline and the
// argument that the result equals the exact top-k under the stated total order.
//
// Device layout (built once by tks_upload_csr, see csr_build.cuh): 6.125 bytes per non-zero
//   val     [nnz]    fp32, as uploaded -- or IEEE half (round to nearest even) in the reference's half-precision
//                     mode (host_spmv_topk_csr_gpu.cu:132-136,151-153): 4.125 bytes per non-zero.  The query is
//                     rounded to half as well; a half x half product is exact in fp32 (11 + 11 significand bits)
//                     and the row sums are accumulated in fp32.
//   col16   [nnz]    u16 : column * 4 = the byte offset of x[col] in shared memory
//                     (columns <= 16383, so 16 bits always suffice: a third less
//                     index traffic than CSR's 32-bit column indices)
//   rowbits [nnz/8]  u8  : bit j of byte i <=> non-zero 8i+j starts a row, so the
//                     kernel never reads row_ptr
//   chunk_start[c], chunk_ord[c] : work units of ~chunk_nnz non-zeros aligned to row
//                     starts; ord = ordinal of the chunk's first row among the
//                     non-empty rows.  row_map[ord] -> row id exists only when the
//                     matrix has empty rows (they can never be candidates: the
//                     reference's COO gold never sees them either).
//
// Kernels per query:  csr_sample_kernel (threshold from a ~1 % sample)
//                  -> csr_topk_main_kernel (the HBM stream, > 95 % of the time)
//                  -> select_topk_kernel   (k best of the surviving candidates)
#pragma once

#include "common.cuh"

namespace tks {

constexpr uint32_t kColOffMask = 0xFFFCu;                   // column * 4
constexpr uint32_t kRowStartBit = 0x80000000u;              // batched kernel's staged column words only
constexpr uint32_t kMaxCols = 16384;                       // exclusive: cols <= 16383 (slot `cols` holds 0.0)
constexpr uint32_t kElemsPerIter = kWarp * 8u;              // 256 non-zeros per warp iteration with fp32 values (512 with 16-bit values)
constexpr uint32_t kMainThreads = 512;
constexpr uint32_t kMainThreadsWide = 576;   // k <= 128 variant: 2 x 18 warps per SM at <= 56 registers per thread
// 16-bit value modes: 16 non-zeros per lane, 72-80 registers, 2 x 12 warps per SM.  (Capped at 64 registers the kernel
// runs 2 x 16 warps without spills but measured slower alone: 0.1708 vs 0.167 ms, r02r.)
constexpr uint32_t kMainThreads16 = 384;
constexpr uint32_t kSampleThreads = 256;
constexpr uint32_t kFull = 0xFFFFFFFFu;

struct CsrDevice {
    const void *val;               // fp32 [nnz], or IEEE half / bfloat16 [nnz] when the kernels are instantiated with VT = 1 / 2
    const uint16_t *col16;         // column * 4
    const uint32_t *col12;         // the same offsets packed to 12 bits each (cols <= 1024 only, else nullptr): what the
                                   // single-query kernels stream when they are instantiated with C12
    const uint8_t *rowbits;        // one row-start bit per non-zero
    const uint64_t *chunk_start;   // n_chunks + 1 entries
    const uint32_t *chunk_ord;     // n_chunks entries
    const uint32_t *row_map;       // ordinal -> row id, or nullptr when every row is non-empty
    uint32_t n_chunks;
    uint32_t cols;
    uint32_t row_offset;           // added to every reported row id
    uint32_t val_type;             // TKS_VALUE_*: 0 fp32, 1 half, 2 bfloat16 (the batched kernel branches on it at run time)
    uint32_t start_align;          // chunk loads start on a multiple of this many non-zeros: 8 / 16 (one lane), 128 with bulk copies
    uint32_t l2_prefetch;          // experiment (TKS_L2PF=n): lane 0 bulk-prefetches the iteration n ahead into L2
};

// Per-query scratch that lives in HBM; zeroed at creation and by the select kernel.
struct RunState {
    uint32_t chunk_counter;   // dynamic scheduler of the main kernel
    uint32_t pool_count;      // candidates appended to the global pool
    uint32_t tau_key;         // ordered-float lower bound on the k-th best score (0 = none)
    uint32_t sample_ticket;   // last-block election of the sample kernel
    uint32_t result_count;    // pool size of the last run (statistics)
    // pipelined submits only (tks_submit): the three kernels of a query run on three streams and hand over through
    // sequence numbers in HBM instead of stream order, so that consecutive queries overlap
    uint32_t tau_seq;         // = seq once the sample kernel has published tau_key
    uint32_t main_ticket;     // CTAs of the main kernel that have handed over their survivors
    uint32_t main_seq;        // = seq once every CTA of the main kernel has
    uint32_t error;           // a main-kernel CTA gave up waiting for the threshold (the query may not have arrived)
    uint32_t pad[7];
};

struct U32x8 { uint32_t w[8]; };
struct U32x4 { uint32_t w[4]; };

// 256-bit streaming load (LDG.E.256 on sm_100): read-only path, no L1 allocation --
// every matrix byte is touched exactly once per query.
__device__ __forceinline__ U32x8 ldg_stream_256(const void *p) {
    U32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]),
                   "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}

__device__ __forceinline__ U32x4 ldg_stream_128(const void *p) {
    U32x4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u8(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }
__device__ __forceinline__ float tau_from_key(uint32_t key) { return key == 0 ? neg_inf() : ordered_to_f32(key); }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// %globaltimer stamps of one pipelined submit (tks_pipeline_stamps): TKS_PIPE_STAMP_WORDS words per query
enum : uint32_t { kStampSampleBegin = 0, kStampSampleEnd = 1, kStampMainBegin = 2, kStampMainEnd = 3,
                  kStampSelectResident = 4, kStampSelectBegin = 5, kStampSelectEnd = 6, kStampWords = 8 };

// One thread waits until *p == want (a sequence number published with st_release_u32 by another grid).  Bounded:
// returns false after timeout_ns of %globaltimer, so a producer that never runs surfaces as an error or a slow path
// instead of a hung device.
__device__ __forceinline__ bool spin_until_eq(const uint32_t *p, uint32_t want, uint64_t timeout_ns) {
    uint64_t t0 = 0;
    for (uint32_t spins = 1;; spins++) {
        if (ld_acquire_u32(p) == want) return true;
        __nanosleep(spins < 64 ? 32 : 256);
        if ((spins & 31u) == 0) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > timeout_ns) return false;
        }
    }
}

// --------------------------------------------------------------------------
// One warp iteration: 32 x EPL consecutive non-zeros, EPL per lane (8 with fp32 values, 16 with 16-bit values:
// both are one 256-bit load of values per lane; with 16 the per-iteration work -- the warp scan, the ballots, the
// loop itself -- is spread over twice as many non-zeros, which is what bounds the 16-bit modes).
//   seg[j]  inclusive segmented sum inside the lane (restarts at row starts)
//   fb      bit j set <=> element j starts a row
//   T       total of the row that ends at the lane's first row start
//   cm      max of seg[j-1] over row starts j >= 1: a superset filter for rows that
//           start and end inside this lane
// All float ops are explicit __fmul_rn/__fadd_rn so that the sample kernel and the
// main kernel produce bit-identical row sums (no FMA contraction either way).
// --------------------------------------------------------------------------
// VT: storage type of the matrix values -- 0 fp32, 1 IEEE half, 2 bfloat16 (TKS_VALUE_*); the 16-bit types are
// widened to fp32 before the multiply.
template <int VT> struct Epl { static constexpr uint32_t v = (VT != 0) ? 16u : 8u; };
constexpr uint32_t elems_per_iter(int vt) { return kWarp * (vt != 0 ? 16u : 8u); }

template <uint32_t EPL>
struct IterState {
    float seg[EPL];
    float T, cm;
    uint32_t fb, nf;
    unsigned fm;   // ballot: lanes with >= 1 row start
};

struct U32x3 { uint32_t w[3]; };
struct U32x6 { uint32_t w[6]; };
template <int VT> struct ValRaw { using type = U32x8; };   // 8 fp32 words or 16 halves
// column offsets of one lane: 16-bit each (8 -> 4 words, 16 -> 8 words) or packed to 12 bits (C12: 3 / 6 words).  Columns
// are at most 1023 whenever the matrix obeys the reference's MAX_COLS = 1024 (types.hpp:55), so column * 4 fits 12 bits
// and the stream shrinks from 6.125 to 5.625 bytes per non-zero (4.125 -> 3.625 with 16-bit values).
template <int VT, bool C12> struct ColRaw { using type = U32x8; };            // 16 x u16
template <> struct ColRaw<0, false> { using type = U32x4; };                  //  8 x u16
template <> struct ColRaw<0, true> { using type = U32x3; };                   //  8 x 12 bits
template <> struct ColRaw<1, true> { using type = U32x6; };                   // 16 x 12 bits
template <> struct ColRaw<2, true> { using type = U32x6; };

template <int VT>
__device__ __forceinline__ typename ValRaw<VT>::type ldg_stream_vals(const void *p) { return ldg_stream_256(p); }
__device__ __forceinline__ uint32_t ldg_stream_32(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
template <int VT, bool C12>
__device__ __forceinline__ typename ColRaw<VT, C12>::type ldg_stream_cols(const void *p) {
    if constexpr (C12 && VT != 0) {
        U32x6 r;   // 24 bytes per lane, 8-byte aligned
#pragma unroll
        for (int i = 0; i < 3; i++)
            asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r.w[2 * i]), "=r"(r.w[2 * i + 1])
                         : "l"(reinterpret_cast<const uint8_t *>(p) + 8 * i));
        return r;
    } else if constexpr (C12) {
        U32x3 r;   // 12 bytes per lane, 4-byte aligned
#pragma unroll
        for (int i = 0; i < 3; i++) r.w[i] = ldg_stream_32(reinterpret_cast<const uint8_t *>(p) + 4 * i);
        return r;
    } else if constexpr (VT != 0) {
        return ldg_stream_256(p);
    } else {
        return ldg_stream_128(p);
    }
}
// column * 4 of element j of the lane
template <bool C12, typename CR>
__device__ __forceinline__ uint32_t col_off_at(const CR &craw, int j) {
    if constexpr (C12) {
        const int bit = 12 * j, wd = bit >> 5, sh = bit & 31;
        if (sh + 12 <= 32) return (sh == 20) ? (craw.w[wd] >> 20) : ((craw.w[wd] >> sh) & 0xFFFu);
        return __funnelshift_r(craw.w[wd], craw.w[wd + 1], sh) & 0xFFFu;
    } else {
        return (j & 1) ? (craw.w[j >> 1] >> 16) : (craw.w[j >> 1] & 0xFFFFu);
    }
}
template <int VT>
__device__ __forceinline__ uint32_t ldg_stream_rowbits(const void *p) {
    if constexpr (VT != 0) {
        uint32_t r;
        asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=r"(r) : "l"(p));
        return r;
    } else {
        return ldg_stream_u8(p);
    }
}

// XSH: the query sits in shared memory as 2^XSH interleaved copies, cell (column, copy) at byte (column * 4 << XSH) +
// copy * 4, and xs_bytes already points at this lane's copy (XSH = 5: one copy per lane, so the gather of a warp's 32
// random columns never meets in a bank; XSH = 0: one copy, ~3.5 wavefronts per gather)
template <bool MASKED, int VT, bool C12, int XSH = 0>
__device__ __forceinline__ void csr_iter(const typename ValRaw<VT>::type &vraw, const typename ColRaw<VT, C12>::type &craw,
                                         uint32_t rbits, const uint8_t *__restrict__ xs_bytes, uint32_t zero_off,
                                         uint32_t lo, uint32_t hi, float carry_in, float &carry_out,
                                         IterState<Epl<VT>::v> &o) {
    constexpr int EPL = (int)Epl<VT>::v;
    const unsigned lane = lane_id();
    float cm = neg_inf();
    // bit j <=> element j starts a row (elements outside [lo, hi) start nothing)
    const uint32_t inmask = ((1u << hi) - 1u) & ~((1u << lo) - 1u);   // elements j in [lo, hi) are inside the chunk
    const uint32_t fb = MASKED ? (rbits & inmask) : rbits;
#pragma unroll
    for (int j = 0; j < EPL; j++) {
        uint32_t c = col_off_at<C12>(craw, j);   // column * 4
        float v;
        if constexpr (VT == 1) {
            const float2 v2 = __half22float2(*reinterpret_cast<const __half2 *>(&vraw.w[j >> 1]));
            v = (j & 1) ? v2.y : v2.x;
        } else if constexpr (VT == 2) {
            // bfloat16 = the upper half of the fp32 word: one shift or mask
            v = __uint_as_float((j & 1) ? (vraw.w[j >> 1] & 0xFFFF0000u) : (vraw.w[j >> 1] << 16));
        } else {
            v = __uint_as_float(vraw.w[j]);
        }
        if (MASKED) {
            const bool in = (inmask >> j) & 1u;
            c = in ? c : zero_off;      // column -> the zero slot behind x
            v = in ? v : 0.0f;
        }
        const float x = *reinterpret_cast<const float *>(xs_bytes + (c << XSH));
        const float p = __fmul_rn(v, x);
        const bool f = (fb >> j) & 1u;
        if (j == 0) {
            o.seg[0] = p;
        } else {
            if (f) cm = fmaxf(cm, o.seg[j - 1]);
            o.seg[j] = f ? p : __fadd_rn(o.seg[j - 1], p);
        }
    }
    o.cm = cm;
    o.fb = fb;
    o.nf = __popc(o.fb);
    o.fm = __ballot_sync(kFull, o.fb != 0);

    // head = sum of the elements before the lane's first row start (whole lane if none): seg[t - 1] for the first
    // row start t >= 1, 0 for t == 0, picked by a select tree over the bits of t
    const uint32_t t = (uint32_t)__ffs((int)o.fb) - 1u;   // first row start (undefined when fb == 0)
    float sel[EPL];
    sel[0] = 0.0f;
#pragma unroll
    for (int j = 1; j < EPL; j++) sel[j] = o.seg[j - 1];
#pragma unroll
    for (int bit = 0; (1 << bit) < EPL; bit++) {
        const bool b = (t >> bit) & 1u;
#pragma unroll
        for (int i = 0; i < (EPL >> (bit + 1)); i++) sel[i] = b ? sel[2 * i + 1] : sel[2 * i];
    }
    const float head = (o.fb == 0) ? o.seg[EPL - 1] : sel[0];

    // inclusive segmented scan of the lane tails; segments restart at lanes with a row start
    const unsigned le = o.fm & lanemask_le();
    const int dist = (int)lane - (le ? (31 - __clz(le)) : 0);
    float I = o.seg[EPL - 1];
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const float up = __shfl_up_sync(kFull, I, dlt);
        if (dist >= dlt) I = __fadd_rn(I, up);
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
    uint64_t *buf;            // this warp's CAP keys in shared memory
    uint32_t cnt;             // warp-uniform
    uint32_t k;
    float tau;                // current lower bound used by the filter
    uint32_t *tau_key_g;      // global lower bound (atomicMax)
    const uint32_t *row_map;
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
            tau = ordered_to_f32(old > tk ? old : tk);
        }
    }
    __device__ __forceinline__ void emit(bool pred, float score, uint32_t ord) {
        const unsigned m = __ballot_sync(kFull, pred);
        if (m == 0) return;
        const uint32_t n = __popc(m);
        if (cnt + n > CAP) compact();
        if (pred) {
            const uint32_t row = (row_map ? row_map[ord] : ord) + row_offset;
            buf[cnt + __popc(m & lanemask_lt())] = make_key(f32_to_ordered(score), row, tie_higher);
        }
        cnt += n;
        __syncwarp();
    }
};

// What a warp carries from iteration to iteration inside a chunk.
struct ChunkCarry {
    uint32_t R;            // ordinal of the row "in progress" (a bogus one before the chunk at first)
    bool first_pending;
    float carry;
};

// One warp iteration's loaded words -> products, segmented sums, candidates.  `edge`: the iteration holds elements
// outside [rel_s, rel_e) (first / last iteration of the chunk), which are neutralised.
template <int VT, bool C12, int XSH = 0, typename Sink>
__device__ __forceinline__ void csr_consume_iter(const typename ValRaw<VT>::type &cv, const typename ColRaw<VT, C12>::type &cc,
                                                 uint32_t cr, bool edge, uint32_t it, int32_t rel_s, int32_t rel_e,
                                                 const uint8_t *__restrict__ xs_bytes, uint32_t zero_off, ChunkCarry &cy,
                                                 Sink &sink) {
    constexpr uint32_t EPL = Epl<VT>::v, EPI = kWarp * EPL;
    const unsigned lane = lane_id();
    IterState<EPL> o;
    float carry_out;
    if (edge) {
        // offsets inside the chunk fit 32 bits (a chunk is far smaller than 2^31 non-zeros)
        const int32_t ebase = (int32_t)(it * EPI + lane * EPL);
        const int32_t l32 = rel_s - ebase, h32 = rel_e - ebase;
        const uint32_t lo = l32 < 0 ? 0u : (l32 > (int32_t)EPL ? EPL : (uint32_t)l32);
        const uint32_t hi = h32 < 0 ? 0u : (h32 > (int32_t)EPL ? EPL : (uint32_t)h32);
        csr_iter<true, VT, C12, XSH>(cv, cc, cr, xs_bytes, zero_off, lo, hi, cy.carry, carry_out, o);
    } else {
        csr_iter<false, VT, C12, XSH>(cv, cc, cr, xs_bytes, zero_off, 0u, EPL, cy.carry, carry_out, o);
    }
    cy.carry = carry_out;

    const bool passT = (o.fb != 0) && (o.T >= sink.tau);
    const unsigned pm = __ballot_sync(kFull, passT || (o.cm >= sink.tau));
    const uint32_t Rtot = __reduce_add_sync(kFull, o.nf);
    if (pm) {
        // ordinal of the row in progress when entering this lane = R + rows started in lower lanes
        uint32_t pre = o.nf;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const uint32_t up = __shfl_up_sync(kFull, pre, dlt);
            if ((int)lane >= dlt) pre += up;
        }
        const uint32_t ord_l = cy.R + (pre - o.nf);
        const bool bogus = cy.first_pending && (lane == (unsigned)(__ffs(o.fm) - 1));
        sink.emit(passT && !bogus, o.T, ord_l);
        if (__any_sync(kFull, o.nf >= 2)) {
            // rows that start AND end inside one lane (length <= EPL - 1)
#pragma unroll
            for (int j = 1; j < (int)EPL; j++) {
                const uint32_t below = o.fb & ((1u << j) - 1u);
                const bool ends_here = ((o.fb >> j) & 1u) && below != 0;
                if (EPL > 8 && !__any_sync(kFull, ends_here)) continue;   // warp-uniform skip
                sink.emit(ends_here && (o.seg[j - 1] >= sink.tau), o.seg[j - 1], ord_l + __popc(below));
            }
        }
    }
    if (o.fm) cy.first_pending = false;
    cy.R += Rtot;
}

// Stream one chunk (or its first max_iters iterations) through `sink`.
template <int VT, bool C12, int XSH = 0, typename Sink>
__device__ __forceinline__ void csr_process_chunk(const CsrDevice &m, const uint8_t *__restrict__ xs_bytes, uint32_t c,
                                                  uint32_t max_iters, Sink &sink) {
    constexpr uint32_t EPL = Epl<VT>::v, EPI = kWarp * EPL;
    const unsigned lane = lane_id();
    const uint64_t s = m.chunk_start[c], e = m.chunk_start[c + 1];
    if (s >= e) return;
    // start of the first load: aligned to one lane's elements (32 bytes of values), or to 128 non-zeros when the matrix is
    // also streamed by the bulk-copy variant (16-byte aligned row-start bits); the sample and the main kernel must agree
    // on it -- the lane a non-zero lands in decides how its row's sum associates
    const uint64_t a0 = s & ~(uint64_t)(m.start_align - 1u);
    const uint64_t n_iter64 = (e - a0 + EPI - 1) / EPI;
    const bool truncated = n_iter64 > max_iters;
    const uint32_t n_iter = truncated ? max_iters : (uint32_t)n_iter64;
    const uint32_t last_iter = (uint32_t)n_iter64 - 1;   // chunks are far smaller than 2^32 * 256 non-zeros
    const int32_t rel_s = (int32_t)(s - a0), rel_e = (int32_t)(e - a0);
    constexpr uint32_t kValBytes = VT != 0 ? 2u : 4u;
    const uint8_t *vp = reinterpret_cast<const uint8_t *>(m.val) + a0 * kValBytes + lane * (EPL * kValBytes);
    // column offsets: 2 bytes each, or 1.5 (a0 is a multiple of 8, so a0 * 3 / 2 is exact and 4-byte aligned)
    constexpr uint32_t kColLaneBytes = C12 ? EPL * 3u / 2u : EPL * 2u;
    const uint8_t *cp = (C12 ? reinterpret_cast<const uint8_t *>(m.col12) + a0 / 2u * 3u
                             : reinterpret_cast<const uint8_t *>(m.col16 + a0)) + lane * kColLaneBytes;
    const uint8_t *rp = m.rowbits + (a0 >> 3) + lane * (EPL / 8u);
    const uint32_t zero_off = m.cols * 4u;

    ChunkCarry cy{m.chunk_ord[c] - 1u, true, 0.0f};

    typename ValRaw<VT>::type nv = ldg_stream_vals<VT>(vp);
    typename ColRaw<VT, C12>::type nc = ldg_stream_cols<VT, C12>(cp);
    uint32_t nr = ldg_stream_rowbits<VT>(rp);
#pragma unroll 2
    for (uint32_t it = 0; it < n_iter; it++) {
        const typename ValRaw<VT>::type cv = nv;
        const typename ColRaw<VT, C12>::type cc = nc;
        const uint32_t cr = nr;
        vp += EPI * kValBytes;
        cp += kWarp * kColLaneBytes;
        rp += EPI / 8u;
#ifdef TKS_EXPERIMENT_L2PF   // measured slower (r02ab: cfg2 main kernel 0.205 -> 0.230-0.234 ms at depth 2..4); compiled out, since
                            // even the untaken branch cost the 16-bit kernel 5 %
        if (m.l2_prefetch && it + m.l2_prefetch < n_iter && lane == 0) {
            // one instruction per array pulls a whole later iteration into L2: no registers, no shared memory, no barrier
            const uint8_t *pv = vp - lane * (EPL * kValBytes) + (size_t)(m.l2_prefetch - 1u) * EPI * kValBytes;
            const uint8_t *pc = cp - lane * kColLaneBytes + (size_t)(m.l2_prefetch - 1u) * kWarp * kColLaneBytes;
            const uint8_t *pr = rp - lane * (EPL / 8u) + (size_t)(m.l2_prefetch - 1u) * (EPI / 8u);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pv), "r"(EPI * kValBytes) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((uintptr_t)pc & ~(uintptr_t)15), "r"((kWarp * kColLaneBytes + 31u) & ~15u) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((uintptr_t)pr & ~(uintptr_t)15), "r"((EPI / 8u + 31u) & ~15u) : "memory");
        }
#endif
        if (it + 1 < n_iter) {
            // lanes whose elements all lie behind the chunk's end (last iteration) fetch nothing: on average half an
            // iteration per chunk, ~3 % of the bytes of a 4096-non-zero chunk
            if ((int32_t)((it + 1u) * EPI + lane * EPL) < rel_e) {
                nv = ldg_stream_vals<VT>(vp); nc = ldg_stream_cols<VT, C12>(cp); nr = ldg_stream_rowbits<VT>(rp);
            } else {
                nr = 0u;   // no row starts; values and columns of this lane are masked by the edge iteration anyway
            }
        }
        csr_consume_iter<VT, C12, XSH>(cv, cc, cr, it == 0 || it == last_iter, it, rel_s, rel_e, xs_bytes, zero_off, cy, sink);
    }
    if (!truncated) {
        // the row in progress at the end of the chunk is complete (chunks end on row boundaries)
        sink.emit(lane == 0 && !cy.first_pending && (cy.carry >= sink.tau), cy.carry, cy.R);
    }
}

// --------------------------------------------------------------------------
// The same stream staged through shared memory by bulk copies (the north-star's "TMA bulk copies and mbarrier
// double-buffering"): every warp owns a ring of kTmaStages stages of one iteration each (values | column offsets |
// row-start bits, three `cp.async.bulk` per stage issued by lane 0, completion counted on the stage's mbarrier), so
// kTmaStages - 1 iterations are in flight per warp without holding them in registers.  Lanes read their 32 bytes of
// values (and, with 16-bit values, of column offsets) as two LDS.128 whose halves are swapped for every second group of
// four lanes, which makes them conflict-free.  Requires chunk loads to start on 128-non-zero boundaries
// (CsrDevice::start_align = 128: `cp.async.bulk` wants 16-byte aligned sources, also for the row-start bits).
// --------------------------------------------------------------------------
constexpr uint32_t kTmaStages = 3;

template <int VT>
struct TmaStage {
    static constexpr uint32_t EPL = Epl<VT>::v, EPI = kWarp * EPL;
    static constexpr uint32_t kValBytes = EPI * (VT != 0 ? 2u : 4u), kColBytes = EPI * 2u, kBitBytes = EPI / 8u;
    static constexpr uint32_t kBytes = kValBytes + kColBytes + kBitBytes;
    static constexpr uint32_t kStride = (kBytes + 127u) & ~127u;
};

struct TmaRing {
    uint32_t base;     // shared address of this warp's kTmaStages stages
    uint32_t bar;      // shared address of this warp's kTmaStages mbarriers
    uint32_t slot;     // stage the next iteration is consumed from
    uint32_t parity;   // its phase
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ U32x4 lds_128(uint32_t addr) {
    U32x4 r;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "r"(addr));
    return r;
}
// a lane's 32 bytes at base + lane * 32, conflict-free: groups of four lanes alternate which half they fetch first
__device__ __forceinline__ U32x8 lds_lane32(uint32_t base, unsigned lane) {
    const uint32_t sw = (lane >> 2) & 1u;
    const uint32_t a = base + lane * 32u;
    const U32x4 f = lds_128(a + sw * 16u), g = lds_128(a + (sw ^ 1u) * 16u);
    U32x8 r;
#pragma unroll
    for (int j = 0; j < 4; j++) { r.w[j] = sw ? g.w[j] : f.w[j]; r.w[4 + j] = sw ? f.w[j] : g.w[j]; }
    return r;
}

template <int VT, typename Sink>
__device__ __forceinline__ void csr_process_chunk_tma(const CsrDevice &m, const uint8_t *__restrict__ xs_bytes, uint32_t c,
                                                      Sink &sink, TmaRing &ring) {
    using S = TmaStage<VT>;
    constexpr uint32_t EPL = S::EPL, EPI = S::EPI;
    const unsigned lane = lane_id();
    const uint64_t s = m.chunk_start[c], e = m.chunk_start[c + 1];
    if (s >= e) return;
    const uint64_t a0 = s & ~127ull;
    const uint32_t n_iter = (uint32_t)((e - a0 + EPI - 1) / EPI);
    const int32_t rel_s = (int32_t)(s - a0), rel_e = (int32_t)(e - a0);
    const uint8_t *gv = reinterpret_cast<const uint8_t *>(m.val) + a0 * (VT != 0 ? 2u : 4u);
    const uint8_t *gc = reinterpret_cast<const uint8_t *>(m.col16 + a0);
    const uint8_t *gr = m.rowbits + (a0 >> 3);
    const uint32_t zero_off = m.cols * 4u;
    ChunkCarry cy{m.chunk_ord[c] - 1u, true, 0.0f};

    // producer side (lane 0): stage of iteration `it` = (ring.slot + it) mod kTmaStages
    uint32_t pslot = ring.slot;
    auto issue = [&](uint32_t it) {
        const uint32_t dst = ring.base + pslot * S::kStride, bar = ring.bar + pslot * 8u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lanes' reads of this stage come first
        mbar_expect_tx(bar, S::kBytes);
        bulk_g2s(dst, gv + (size_t)it * S::kValBytes, S::kValBytes, bar);
        bulk_g2s(dst + S::kValBytes, gc + (size_t)it * S::kColBytes, S::kColBytes, bar);
        bulk_g2s(dst + S::kValBytes + S::kColBytes, gr + (size_t)it * S::kBitBytes, S::kBitBytes, bar);
        pslot = (pslot + 1u == kTmaStages) ? 0u : pslot + 1u;
    };
    if (lane == 0) {
        const uint32_t pre = n_iter < kTmaStages - 1u ? n_iter : kTmaStages - 1u;
        for (uint32_t i = 0; i < pre; i++) issue(i);
    }
    for (uint32_t it = 0; it < n_iter; it++) {
        // refill the stage the previous iteration was consumed from (every lane has read it: __syncwarp below)
        if (lane == 0 && it + kTmaStages - 1u < n_iter) issue(it + kTmaStages - 1u);
        const uint32_t st = ring.base + ring.slot * S::kStride, bar = ring.bar + ring.slot * 8u;
        uint32_t tries = 0;
        while (!mbar_try_wait(bar, ring.parity)) {
            if (++tries > (1u << 22)) __trap();   // a copy that never lands must not hang the device
        }
        typename ValRaw<VT>::type cv = lds_lane32(st, lane);
        typename ColRaw<VT, false>::type cc;
        uint32_t cr;
        if constexpr (VT != 0) {
            cc = lds_lane32(st + S::kValBytes, lane);
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(cr) : "r"(st + S::kValBytes + S::kColBytes + lane * 2u));
        } else {
            cc = lds_128(st + S::kValBytes + lane * 16u);
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(cr) : "r"(st + S::kValBytes + S::kColBytes + lane));
        }
        __syncwarp();
        if (++ring.slot == kTmaStages) { ring.slot = 0; ring.parity ^= 1u; }
        csr_consume_iter<VT, false>(cv, cc, cr, it == 0 || it + 1 == n_iter, it, rel_s, rel_e, xs_bytes, zero_off, cy, sink);
    }
    sink.emit(lane == 0 && !cy.first_pending && (cy.carry >= sink.tau), cy.carry, cy.R);
}

// --------------------------------------------------------------------------
// Exact top-k threshold of n keys (load(i), normally shared memory) by a whole
// CTA: MSB radix select, 8 bits per pass, warp-aggregated histogram updates (the
// keys of one pass mostly share their digit; plain atomics would serialise).
// Returns a threshold `thr` such that exactly-or-at-least k keys are >= thr and
// every key >= thr is among the k largest when keys are unique; stops as soon
// as a whole bucket is needed (usually after 3-4 passes).  Returns 0 when n < k.
// hist: 256 words of shared memory; s_bin/s_above: 2 words.
// --------------------------------------------------------------------------
template <typename KeyT, bool EARLY_EXIT, typename LoadF>
__device__ __forceinline__ KeyT block_radix_select(LoadF load, uint32_t n, uint32_t k, uint32_t *hist,
                                                   uint32_t *s_bin, uint32_t *s_above) {
    // hist: one private 256-bin histogram per warp (blockDim/32 * 256 words): same-digit updates only
    // contend inside a warp, where shared-memory atomics are cheap; bins are summed across warps after.
    if (n < k) return (KeyT)0;
    const uint32_t nwarps = blockDim.x / kWarp;
    uint32_t *my = hist + (threadIdx.x / kWarp) * 256u;
    KeyT prefix = 0, pmask = 0;
    uint32_t need = k;
    for (int shift = (int)sizeof(KeyT) * 8 - 8; shift >= 0; shift -= 8) {
        for (uint32_t i = threadIdx.x; i < nwarps * 256u; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const KeyT key = load(i);
            if ((key & pmask) == prefix) atomicAdd(&my[(uint32_t)(key >> shift) & 0xFFu], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t tot = 0;
            for (uint32_t w = 0; w < nwarps; w++) tot += hist[w * 256u + threadIdx.x];
            hist[threadIdx.x] = tot;   // thread t only ever touches bin t of every slice: no race
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // suffix sums over 256 bins by one warp: lane l owns bins 8l..8l+7
            uint32_t loc[8], tot = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) { loc[b] = hist[threadIdx.x * 8 + b]; tot += loc[b]; }
            uint32_t suf = tot;   // inclusive suffix over lanes
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const uint32_t dn = __shfl_down_sync(kFull, suf, dlt);
                if (threadIdx.x + dlt < 32) suf += dn;
            }
            uint32_t above = suf - tot;   // keys in bins of higher lanes
            // the wanted bin is the highest b with (keys in bins > b) < need <= (keys in bins >= b)
#pragma unroll
            for (int b = 7; b >= 0; b--) {
                if (above < need && above + loc[b] >= need) {
                    s_bin[0] = threadIdx.x * 8 + b;
                    s_above[0] = above;
                    s_above[1] = loc[b];
                }
                above += loc[b];
            }
        }
        __syncthreads();
        prefix |= (KeyT)s_bin[0] << shift;
        pmask |= (KeyT)0xFF << shift;
        need -= s_above[0];
        const bool whole_bucket = (need == s_above[1]);
        __syncthreads();
        if (EARLY_EXIT && whole_bucket) break;   // every key of this bucket is wanted: prefix (low bits 0) is the threshold
    }
    return prefix;
}

// --------------------------------------------------------------------------
// A lower bound on the k-th largest of n 32-bit keys (key 0 = absent) in ONE histogram pass: 2048 bins laid
// linearly over [min, max] of the keys present, so that the bins are evenly filled whatever bits the keys share
// (the ordered-float scores of one query agree in sign and most of the exponent; an MSB radix digit would put
// them all in two or three bins).  Returns the lower edge `thr` of the bin that holds the k-th largest key:
// at least k keys are >= thr, and at most (k - 1 + that bin's count) are.  Returns 0 when fewer than k keys
// are present.  Every thread of the CTA calls it and gets the same value.
// scratch: kHistScratchWords words of shared memory.
// --------------------------------------------------------------------------
constexpr uint32_t kHistBins = 2048;
constexpr uint32_t kHistGroups = kHistBins / 32;
constexpr uint32_t kHistScratchWords = kHistBins + kHistGroups + 3 * 32 + 4;

template <typename LoadF>
__device__ __forceinline__ uint32_t block_hist_threshold(LoadF load32, uint32_t n, uint32_t k, uint32_t *scratch,
                                                         uint32_t *bin_count_out = nullptr) {
    uint32_t *hist = scratch, *gs = scratch + kHistBins, *red = gs + kHistGroups, *res = red + 3 * 32;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x, nwarps = nthr / kWarp;
    const unsigned lane = lane_id();
    uint32_t lo = 0xFFFFFFFFu, hi = 0u, cnt = 0u;
    for (uint32_t i = tid; i < n; i += nthr) {
        const uint32_t key = load32(i);
        if (key) { lo = min(lo, key); hi = max(hi, key); cnt++; }
    }
    lo = __reduce_min_sync(kFull, lo);
    hi = __reduce_max_sync(kFull, hi);
    cnt = __reduce_add_sync(kFull, cnt);
    __syncthreads();   // scratch may still be read by the caller's previous use
    if (lane == 0) { red[tid / kWarp] = lo; red[32 + tid / kWarp] = hi; red[64 + tid / kWarp] = cnt; }
    for (uint32_t i = tid; i < kHistBins; i += nthr) hist[i] = 0u;
    __syncthreads();
    lo = 0xFFFFFFFFu; hi = 0u; cnt = 0u;
    for (uint32_t w = 0; w < nwarps; w++) { lo = min(lo, red[w]); hi = max(hi, red[32 + w]); cnt += red[64 + w]; }
    if (cnt < k) return 0u;   // uniform: every thread computed the same cnt
    const uint32_t range = hi - lo;
    const uint32_t sh = range < kHistBins ? 0u : (32u - (uint32_t)__clz(range)) - 11u;   // (range >> sh) < 2048
    for (uint32_t i = tid; i < n; i += nthr) {
        const uint32_t key = load32(i);
        if (key) atomicAdd(&hist[(key - lo) >> sh], 1u);
    }
    __syncthreads();
    if (tid < kHistGroups) {   // group t = bins [32t, 32t + 32), read skewed: conflict-free
        uint32_t sum = 0;
#pragma unroll 8
        for (uint32_t i = 0; i < 32; i++) sum += hist[tid * 32u + ((i + tid) & 31u)];
        gs[tid] = sum;
    }
    __syncthreads();
    if (tid < kWarp) {
        // lane l owns groups 2l (lower) and 2l + 1 (higher); suffix sums over lanes
        const uint32_t g0 = gs[2 * lane], g1 = gs[2 * lane + 1], tot = g0 + g1;
        uint32_t suf = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t dn = __shfl_down_sync(kFull, suf, d);
            if (lane + d < 32) suf += dn;
        }
        const uint32_t above = suf - tot;   // keys in groups of higher lanes
        int grp = -1;
        uint32_t above_g = 0;
        if (above < k && above + g1 >= k) { grp = 2 * (int)lane + 1; above_g = above; }
        else if (above + g1 < k && above + tot >= k) { grp = 2 * (int)lane; above_g = above + g1; }
        const unsigned who = __ballot_sync(kFull, grp >= 0);
        const int src = __ffs(who) - 1;   // exactly one lane (cnt >= k)
        grp = __shfl_sync(kFull, grp, src);
        above_g = __shfl_sync(kFull, above_g, src);
        const uint32_t c = hist[(uint32_t)grp * 32u + lane];
        uint32_t sufb = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t dn = __shfl_down_sync(kFull, sufb, d);
            if (lane + d < 32) sufb += dn;
        }
        const uint32_t ab = above_g + sufb - c;   // keys in higher bins
        if (ab < k && ab + c >= k) {
            res[0] = lo + (((uint32_t)grp * 32u + lane) << sh);
            res[1] = ab + c;                      // keys >= the returned threshold
        }
    }
    __syncthreads();
    if (bin_count_out) *bin_count_out = res[1];
    return res[0];
}

// --------------------------------------------------------------------------
// Kernel 1: the sample.  Warp w reduces the first sample_iters iterations (256 non-zeros each; ~1 % of the
// matrix in total) of chunk w * n_chunks / n_sample exactly as the main kernel will and publishes its best completed row.
// The k-th largest of these warp maxima is the score of k distinct real rows, hence a valid lower bound on
// the k-th best score; the last CTA to finish derives it in one histogram pass (block_hist_threshold: the lower
// edge of the bin that holds the k-th largest maximum -- still a lower bound) and publishes it in st->tau_key.
// Dynamic shared memory: max((cols+1)*4, kHistScratchWords*4) bytes.
// --------------------------------------------------------------------------
// The query as the kernels see it: rounded to the storage type of the values and widened again.
template <int VT>
__device__ __forceinline__ float query_value(float x) {
    if constexpr (VT == 1) return __half2float(__float2half_rn(x));
    else if constexpr (VT == 2) return __bfloat162float(__float2bfloat16_rn(x));
    else return x;
}

// <= 64 registers (4 CTAs of 256 threads per SM): its warps must fit into the register holes two main-kernel CTAs leave
// in every SM sub-partition when it runs beside them (pipelined submits)
template <int VT, bool C12 = false>
__global__ void __launch_bounds__(kSampleThreads, 4) csr_sample_kernel(CsrDevice m, const float *__restrict__ x,
                                                                     RunState *st, uint32_t *sample_keys,
                                                                     uint32_t n_sample,
                                                                     uint32_t sample_iters, uint32_t k, uint32_t seq,
                                                                     uint64_t *stamp) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    pdl_trigger();   // the main kernel's CTAs may take the SMs this grid leaves and stage the query meanwhile
    if (stamp && blockIdx.x == 0 && threadIdx.x == 0) stamp[kStampSampleBegin] = global_timer_ns();
    float *xs = reinterpret_cast<float *>(smem_raw);
    for (uint32_t i = threadIdx.x; i <= m.cols; i += blockDim.x) xs[i] = (i < m.cols) ? query_value<VT>(x[i]) : 0.0f;
    __syncthreads();
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    if (gw < n_sample) {
        MaxSink sink{0.0f, false, neg_inf()};
        // sampled work units are spread evenly over the whole stream, also when units / samples is not a whole number
        const uint32_t c = (uint32_t)(((uint64_t)gw * m.n_chunks) / n_sample);
        if (c < m.n_chunks) csr_process_chunk<VT, C12>(m, smem_raw, c, sample_iters, sink);
        uint32_t key = sink.any ? f32_to_ordered(sink.best) : 0u;
        key = __reduce_max_sync(kFull, key);
        if (lane_id() == 0) sample_keys[gw] = key;
    }
    // last block: threshold from all the maxima
    __shared__ uint32_t s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&st->sample_ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    // x is no longer needed by this block: its shared memory becomes the histogram scratch; the maxima are read from
    // L2 (three passes over <= 32 KB), so that a sample CTA needs only ~9 KB and fits beside two main-kernel CTAs
    __syncthreads();
    const uint32_t thr = block_hist_threshold([&](uint32_t i) { return __ldcg(sample_keys + i); }, n_sample, k,
                                              reinterpret_cast<uint32_t *>(smem_raw));
    if (threadIdx.x == 0) {
        if (thr != 0) atomicMax(&st->tau_key, thr);
        st->sample_ticket = 0;
        if (stamp) stamp[kStampSampleEnd] = global_timer_ns();
        if (seq) {   // pipelined submit: the main kernel of this query waits for this number, not for stream order
            __threadfence();
            st_release_u32(&st->tau_seq, seq);
        }
    }
}

// --------------------------------------------------------------------------
// Kernel 2: the stream.  Persistent CTAs; every warp pulls chunks from a global
// counter, reduces them, and keeps rows with score >= tau in a private
// shared-memory buffer (sorted and cut to k only if it ever fills).
// --------------------------------------------------------------------------
// TMA: the stream arrives through per-warp rings of bulk copies (csr_process_chunk_tma) instead of register loads.
// Shared memory then holds, behind the query and the candidate buffers, warps x kTmaStages stages (128-byte aligned)
// and warps x kTmaStages mbarriers: main_tma_extra_smem().
template <int VT>
__host__ __device__ constexpr size_t main_tma_extra_smem(uint32_t warps) {
    return 128u + (size_t)warps * kTmaStages * (TmaStage<VT>::kStride + 8u);
}

// XSH = 5 (TKS_XCOPIES=1, measurement variant): one CTA per SM whose 32 query copies (128 bytes per column) make the
// gather conflict-free; CTA size kMainThreadsX / kMainThreadsX16.
constexpr uint32_t kMainThreadsX = 1024, kMainThreadsX16 = 768;
template <int CAP, int VT, bool TMA = false, bool C12 = false, int XSH = 0>
__global__ void __launch_bounds__(XSH ? (VT != 0 ? kMainThreadsX16 : kMainThreadsX)
                                      : (VT != 0 ? kMainThreads16 : (CAP == 256 ? kMainThreadsWide : kMainThreads)), XSH ? 1 : 2)
csr_topk_main_kernel(CsrDevice m, const float *__restrict__ x, RunState *st, uint64_t *pool, uint32_t k,
                     int tie_higher, uint32_t seq, uint32_t tau_wait_us, uint64_t *stamp) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float *xs = reinterpret_cast<float *>(smem_raw);
    uint64_t *bufs = reinterpret_cast<uint64_t *>(smem_raw + ((((m.cols + 1u) * 4u << XSH) + 15u) & ~15u));
    // cell i of the staged query: column i >> XSH (every copy holds the same value)
    const uint32_t n_cells = (m.cols + 1u) << XSH;
    TmaRing ring{0u, 0u, 0u, 0u};
    if constexpr (TMA) {
        const uint32_t warps = blockDim.x / kWarp, w = threadIdx.x / kWarp;
        const uint32_t after_bufs = (uint32_t)__cvta_generic_to_shared(bufs + (size_t)warps * CAP);
        const uint32_t ring0 = (after_bufs + 127u) & ~127u;
        ring.base = ring0 + w * kTmaStages * TmaStage<VT>::kStride;
        ring.bar = ring0 + warps * kTmaStages * TmaStage<VT>::kStride + w * kTmaStages * 8u;
        if (lane_id() == 0) {
            for (uint32_t i = 0; i < kTmaStages; i++) mbar_init(ring.bar + i * 8u, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    pdl_trigger();   // the select kernel's CTA may be set up while this grid drains
    if (seq == 0) {
        for (uint32_t i = threadIdx.x; i < n_cells; i += blockDim.x) xs[i] = ((i >> XSH) < m.cols) ? query_value<VT>(x[i >> XSH]) : 0.0f;
        __syncthreads();
        pdl_wait();  // the query was complete before the sample kernel started; tau and the counters are not
    } else {
        // Pipelined submit (tks_submit): this grid follows the PREVIOUS query's main kernel in its stream and shares
        // nothing with it (per-slot state), so it does not wait for it -- its CTAs start streaming as that grid's CTAs
        // retire, and consecutive queries form one continuous stream.  What it needs is this query's threshold, which
        // the sample kernel published from another stream (normally long ago); that kernel ran after the query's bytes
        // had arrived (tks_submit_host copies them on the sample stream), so the query is staged only now.  Should the
        // threshold never arrive in time, the stream runs unfiltered and the query is reported as failed (RunState::error).
        if (threadIdx.x == 0) {
            if (!spin_until_eq(&st->tau_seq, seq, (uint64_t)tau_wait_us * 1000ull)) atomicOr(&st->error, 1u);
            if (stamp && blockIdx.x == 0) stamp[kStampMainBegin] = global_timer_ns();
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_cells; i += blockDim.x) xs[i] = ((i >> XSH) < m.cols) ? query_value<VT>(__ldcg(x + (i >> XSH))) : 0.0f;
        __syncthreads();
    }

    const unsigned lane = lane_id();
    const uint8_t *xs_lane = smem_raw + (XSH ? lane * 4u : 0u);   // this lane's copy
    PoolSink<CAP> sink;
    sink.buf = bufs + (threadIdx.x / kWarp) * CAP;
    sink.cnt = 0;
    sink.tau = neg_inf();
    sink.k = k;
    sink.tau_key_g = &st->tau_key;
    sink.row_map = m.row_map;
    sink.row_offset = m.row_offset;
    sink.tie_higher = tie_higher;

    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(&st->chunk_counter, 1u);
        c = __shfl_sync(kFull, c, 0);
        if (c >= m.n_chunks) break;
        sink.tau = fmaxf(sink.tau, tau_from_key(ld_relaxed_u32(&st->tau_key)));
        if constexpr (TMA) csr_process_chunk_tma<VT>(m, smem_raw, c, sink, ring);
        else csr_process_chunk<VT, C12, XSH>(m, xs_lane, c, 0xFFFFFFFFu, sink);
    }

    // hand the survivors to the global pool (filtered by the freshest bound)
    const uint32_t tk = ld_relaxed_u32(&st->tau_key);
    __syncwarp();
    for (uint32_t base = 0; base < sink.cnt; base += kWarp) {
        const uint32_t i = base + lane;
        const uint64_t key = (i < sink.cnt) ? sink.buf[i] : 0ull;
        const bool keep = (i < sink.cnt) && (key_score(key) >= tk);
        const unsigned mk = __ballot_sync(kFull, keep);
        if (mk) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(&st->pool_count, (uint32_t)__popc(mk));
            pos = __shfl_sync(kFull, pos, 0);
            if (keep) pool[pos + __popc(mk & lanemask_lt())] = key;
        }
    }
    if (seq) {
        // the last CTA to hand over its survivors tells the select kernel (another stream) that the pool is complete
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&st->main_ticket, 1u) == gridDim.x - 1) {
            if (stamp) stamp[kStampMainEnd] = global_timer_ns();
            __threadfence();
            st_release_u32(&st->main_seq, seq);
        }
    }
}

// ---- constants of the select kernel (Kernel 3) ----
constexpr uint32_t kSelectThreads = 1024;
constexpr uint32_t kSelectSortCap = 2048;     // keys sorted directly (16 KB static)
constexpr uint32_t kSelectRankSortMax = 512;  // up to here by rank (one pass, no barriers), above by a bitonic network
constexpr uint32_t kSelectSmemKeys = 16384;   // pool keys staged in dynamic shared memory (128 KB)
constexpr uint32_t kSelectDynSmem = (kSelectThreads / kWarp) * 1024u + kSelectSmemKeys * 8u;
constexpr uint32_t kSelectLeanThreads = 512;   // pipelined submits: no staged keys, per-warp histograms only
constexpr uint32_t kSelectLeanDynSmem = (kSelectLeanThreads / kWarp) * 1024u;

// ---- candidate exchange over peer memory: shared pieces (the protocol is described at Kernel 4 below) ----
constexpr uint32_t kPeerMaxWorld = 8;
constexpr uint32_t kPeerTimeout = 0xFFFFFFFEu;

struct PeerExchange {
    uint64_t *window[kPeerMaxWorld];   // window[r]: rank r's window as mapped in THIS process (window[rank] = local)
    uint32_t world, rank, kmax;
    uint32_t timeout_ms;               // bound of the wait for the peers' records (TKS_SPIN_TIMEOUT_MS, default 2000)
};

__host__ __device__ constexpr size_t peer_window_bytes(uint32_t kmax) { return 2ull * kPeerMaxWorld * kmax * 2ull * sizeof(uint64_t); }

__device__ __forceinline__ void st_volatile_v2_u64(uint64_t *p, uint64_t a, uint64_t b) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_volatile_v2_u64(const uint64_t *p, uint64_t &a, uint64_t &b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// This rank's key of output slot `slot` (0 = absent) into every rank's window, each word tagged with the step.
__device__ __forceinline__ void peer_push_key(const PeerExchange &px, uint32_t seq, uint32_t slot, uint64_t key) {
    const uint64_t tag = (uint64_t)seq << 32;
    const size_t rec = ((((size_t)(seq & 1u)) * px.world + px.rank) * px.kmax + slot) * 2u;
    for (uint32_t d = 0; d < px.world; d++) {
        const uint32_t p = (px.rank + 1u + d) % px.world;   // peers first, the local copy last
        st_volatile_v2_u64(px.window[p] + rec, tag | (key >> 32), tag | (key & 0xFFFFFFFFull));
    }
}

// Whole CTA: wait until every rank's k records of this step sit in the local window, gather them into `keys`
// (shared memory, kSelectSortCap entries), merge and write the global top-k.  *s_timeout / *s_present: shared words
// zeroed by the caller before a barrier.
__device__ __forceinline__ void peer_poll_and_merge(const PeerExchange &px, uint32_t seq, uint32_t k, int tie_higher,
                                                    uint64_t *keys, uint32_t *s_timeout, uint32_t *s_present,
                                                    uint64_t *res_keys, uint32_t *res_idx, float *res_val,
                                                    uint32_t *res_count) {
    const uint32_t tid = threadIdx.x;
    const uint32_t parity = seq & 1u;
    const uint32_t n = px.world * k;   // <= kSelectSortCap (checked by the host)
    uint32_t present = 0;
    for (uint32_t t = tid; t < n; t += blockDim.x) {
        const uint32_t r = t / k, i = t - r * k;
        const uint64_t *src = px.window[px.rank] + (((size_t)parity * px.world + r) * px.kmax + i) * 2u;
        uint64_t a, b;
        uint64_t t0 = 0;
        uint32_t spins = 0;
        for (;;) {
            ld_volatile_v2_u64(src, a, b);
            if ((a >> 32) == seq && (b >> 32) == seq) break;
            if ((++spins & 63u) == 0) {
                const uint64_t now = global_timer_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > (uint64_t)px.timeout_ms * 1000000ull || *reinterpret_cast<volatile uint32_t *>(s_timeout)) { *s_timeout = 1; break; }
            }
        }
        const uint64_t key = (a << 32) | (b & 0xFFFFFFFFull);
        keys[t] = key;
        present += key != 0ull;
    }
    present = __reduce_add_sync(kFull, present);
    if (lane_id() == 0 && present) atomicAdd(s_present, present);
    __syncthreads();
    if (*s_timeout) {
        if (tid == 0) *res_count = kPeerTimeout;
        return;
    }
    const uint32_t cnt = *s_present < k ? *s_present : k;
    // Merge of `world` lists that are each sorted descending (absent = 0 at the tail): the output slot of a key is its
    // position in its own list plus, for every other list, the number of keys ahead of it there -- one binary search
    // per list (world x log2(k) probes per key instead of world x k comparisons).  Equal keys (only zeros, or the same
    // row reported twice) are ordered by rank so that slots stay unique.
    for (uint32_t t = tid; t < n; t += blockDim.x) {
        const uint32_t r = t / k, i = t - r * k;
        const uint64_t mine = keys[t];
        uint32_t slot = i;
        for (uint32_t o = 0; o < px.world; o++) {
            if (o == r) continue;
            const uint64_t *lst = keys + (size_t)o * k;
            uint32_t lo = 0, hi = k;   // first index whose key is not ahead of `mine`
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                const uint64_t other = lst[mid];
                const bool ahead = (other > mine) || (other == mine && o < r);
                if (ahead) lo = mid + 1; else hi = mid;
            }
            slot += lo;
        }
        if (slot < cnt) {
            res_keys[slot] = mine;
            res_idx[slot] = key_row(mine, tie_higher);
            res_val[slot] = ordered_to_f32(key_score(mine));
        }
    }
    for (uint32_t i = cnt + tid; i < k; i += blockDim.x) { res_keys[i] = 0ull; res_idx[i] = 0u; res_val[i] = 0.0f; }
    if (tid == 0) *res_count = cnt;
}

// --------------------------------------------------------------------------
// Kernel 3: k best keys of a pool, one CTA.  The pool is staged in shared memory
// (up to kSelectSmemKeys keys; larger pools are read from L2 on every pass),
// reduced with the radix select to the <= kSelectSortCap keys above the
// threshold, and those are sorted.  Also resets the per-query scratch.
// Dynamic shared memory: kSelectDynSmem bytes (per-warp histograms + staged keys).
// --------------------------------------------------------------------------

constexpr uint32_t kPoolOverflow = 0xFFFFFFFFu;   // *out_count when a capped pool overflowed (batched mode)

// The select kernel leaves the per-query scratch ready for the next run (one thread).
__device__ __forceinline__ void reset_run_state(RunState *st, uint32_t pool_size) {
    st->result_count = pool_size;   // for tks_get_stats
    st->chunk_counter = 0;
    st->pool_count = 0;
    st->tau_key = 0;
    st->sample_ticket = 0;
    st->main_ticket = 0;
}

// Grid: one CTA per query.  CTA q reads pool + q * pool_stride; its key count is st[q].pool_count (st != nullptr)
// or pool_count_imm; results go to out_* + q * out_stride and out_count[q].  pool_cap != 0: a count above it
// means keys were dropped -> out_count = kPoolOverflow and nothing else is written.
// EXCHANGE (several GPUs, one query): the k sorted keys do not go to out_* but straight into every rank's peer window
// (peer_push_key), and the same CTA then waits for the other ranks' lists and writes the GLOBAL top-k to out_*
// (peer_poll_and_merge) -- local select, exchange and merge in one launch.
template <bool EXCHANGE>
__global__ void __launch_bounds__(kSelectThreads, 2)   // <= 32 registers: the CTA of a pipelined submit waits beside a main-kernel CTA
select_topk_kernel(const uint64_t *__restrict__ pool, uint32_t pool_stride, RunState *st, uint32_t pool_count_imm,
                   uint32_t pool_cap, uint32_t k, int tie_higher, uint64_t *out_keys, uint32_t *out_idx,
                   float *out_val, uint32_t out_stride, uint32_t *out_count, uint32_t *pass_counter,
                   PeerExchange px, uint32_t seq, uint32_t main_wait_seq, uint32_t main_wait_ms, uint64_t *stamp,
                   uint32_t stage_keys) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);                                  // 32 KB
    // stage_keys: keys the dynamic shared memory holds behind the histograms (kSelectSmemKeys in the ordinary launch;
    // 0 in the lean launch of a pipelined submit, whose CTA has to fit beside a main-kernel CTA: 512 threads, 16 KB)
    uint64_t *staged = reinterpret_cast<uint64_t *>(smem_raw + (blockDim.x / kWarp) * 1024u);
    __shared__ uint64_t keys[kSelectSortCap];
    __shared__ uint32_t s_bin, s_above[2], s_cnt, s_timeout, s_present;
    const uint32_t tid = threadIdx.x;
    const uint32_t q = blockIdx.x;
    if (EXCHANGE && tid == 0) { s_timeout = 0; s_present = 0; }
    RunState *st_reset = st ? st + q : nullptr;
    pool += (size_t)q * pool_stride;
    out_keys += (size_t)q * out_stride;
    out_idx += (size_t)q * out_stride;
    out_val += (size_t)q * out_stride;
    out_count += q;
    pdl_wait();      // the pool and its count belong to the kernel before this one
    if (main_wait_seq) {
        // pipelined submit: this CTA was launched on its own stream, possibly long before the main kernel of its
        // query finished; the last CTA of that grid publishes main_seq (csr_topk_main_kernel)
        __shared__ uint32_t s_main_ok;
        if (tid == 0) {
            if (stamp) stamp[kStampSelectResident] = global_timer_ns();
            s_main_ok = spin_until_eq(&st_reset->main_seq, main_wait_seq, (uint64_t)main_wait_ms * 1000000ull) ? 1u : 0u;
            if (stamp) stamp[kStampSelectBegin] = global_timer_ns();
        }
        __syncthreads();
        if (s_main_ok && st_reset->error) {
            // the stream ran without its threshold, i.e. the sample kernel (and with it, possibly, the query's copy)
            // never completed in time: the result cannot be trusted
            __syncthreads();
            if (tid == 0) { st_reset->error = 0; s_main_ok = 2; }
            __syncthreads();
        }
        if (s_main_ok != 1) {
            if (tid == 0) {
                *out_count = kPeerTimeout;
                if (s_main_ok == 2) reset_run_state(st_reset, st_reset->pool_count);   // the main kernel is done: the slot is reusable
            }
            return;
        }
    }
    const uint32_t n = st_reset ? st_reset->pool_count : pool_count_imm;
    if (pool_cap != 0 && n > pool_cap) {
        __syncthreads();   // everyone has read pool_count
        if (tid == 0) {
            *out_count = kPoolOverflow;
            reset_run_state(st_reset, n);
            if (pass_counter && (q % 32u) == 0) pass_counter[q / 32u] = 0;
        }
        return;
    }

    uint32_t m = 0;   // number of keys in keys[]
    if (n <= kSelectRankSortMax) {
        for (uint32_t i = tid; i < n; i += blockDim.x) keys[i] = __ldcg(pool + i);
        m = n;
    } else {
        const bool in_smem = n <= stage_keys;
        if (in_smem) {
            for (uint32_t i = tid; i < n; i += blockDim.x) staged[i] = __ldcg(pool + i);
            __syncthreads();
        }
        // the pool was written by another grid, possibly while this CTA was already resident: L2, never L1
        auto load = [&](uint32_t i) { return in_smem ? staged[i] : __ldcg(pool + i); };
        // one histogram pass over the score halves: every key whose score reaches the bin of the k-th best score
        // is kept (at least k keys, a few more), the exact order is settled by the sort below
        uint32_t reach = 0;
        const uint32_t thr32 = block_hist_threshold([&](uint32_t i) { return (uint32_t)(load(i) >> 32); }, n, k, hist, &reach);
        uint64_t thr = (uint64_t)thr32 << 32;
        if (thr32 == 0 || reach > kSelectSortCap) {
            // fewer than k keys, or a bin crowded with (nearly) equal scores: the exact radix select on the full keys
            thr = block_radix_select<uint64_t, true>(load, n, k, hist, &s_bin, s_above);
        }
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += blockDim.x) {
            const uint64_t key = load(i);
            if (key >= thr && key != 0ull) {
                const uint32_t pos = atomicAdd(&s_cnt, 1u);
                if (pos < kSelectSortCap) keys[pos] = key;
            }
        }
        __syncthreads();
        m = s_cnt < kSelectSortCap ? s_cnt : kSelectSortCap;
    }
    const uint32_t cnt = m < k ? m : k;
    __syncthreads();
    if (m <= kSelectRankSortMax) {
        // rank sort: four adjacent lanes count the keys above key i (keys are unique: they embed the row id;
        // equal keys would still get distinct ranks through the position tie-break), the rank is the output slot
        const uint32_t items = (4u * m + kWarp - 1u) & ~(kWarp - 1u);
        for (uint32_t w = tid; w < items; w += blockDim.x) {
            const uint32_t i = w >> 2, part = w & 3u;
            const uint64_t mine = (i < m) ? keys[i] : 0ull;
            uint32_t r = 0;
            for (uint32_t j = part; j < m; j += 4u) {
                const uint64_t other = keys[j];
                r += (other > mine) || (other == mine && j < i);
            }
            r += __shfl_xor_sync(kFull, r, 1);
            r += __shfl_xor_sync(kFull, r, 2);
            if (part == 0 && i < m && r < k) {
                if (EXCHANGE) {
                    peer_push_key(px, seq, r, mine);
                } else {
                    out_keys[r] = mine;
                    out_idx[r] = key_row(mine, tie_higher);
                    out_val[r] = ordered_to_f32(key_score(mine));
                }
            }
        }
        for (uint32_t i = cnt + tid; i < k; i += blockDim.x) {
            if (EXCHANGE) peer_push_key(px, seq, i, 0ull);
            else { out_keys[i] = 0ull; out_idx[i] = 0u; out_val[i] = 0.0f; }
        }
    } else {
        uint32_t n2 = 32;
        while (n2 < m) n2 <<= 1;
        for (uint32_t i = m + tid; i < n2; i += blockDim.x) keys[i] = 0ull;
        bitonic_sort_desc(keys, n2, tid, blockDim.x, [] { __syncthreads(); });
        for (uint32_t i = tid; i < k; i += blockDim.x) {
            const uint64_t key = (i < cnt) ? keys[i] : 0ull;
            if (EXCHANGE) {
                peer_push_key(px, seq, i, key);
            } else {
                out_keys[i] = key;
                out_idx[i] = (i < cnt) ? key_row(key, tie_higher) : 0u;
                out_val[i] = (i < cnt) ? ordered_to_f32(key_score(key)) : 0.0f;
            }
        }
    }
    if (EXCHANGE) {
        __syncthreads();   // every thread is done with keys[] (the rank sort reads it) before the gather overwrites it
        peer_poll_and_merge(px, seq, k, tie_higher, keys, &s_timeout, &s_present, out_keys, out_idx, out_val, out_count);
        if (tid == 0) {
            if (st_reset) reset_run_state(st_reset, n);
            if (stamp) stamp[kStampSelectEnd] = global_timer_ns();
        }
        return;
    }
    if (tid == 0) {
        *out_count = cnt;
        if (st_reset) reset_run_state(st_reset, n);
        if (pass_counter && (q % 32u) == 0) pass_counter[q / 32u] = 0;
        if (stamp) stamp[kStampSelectEnd] = global_timer_ns();
    }
}

// --------------------------------------------------------------------------
// Kernel 4 (several GPUs): candidate exchange over peer memory + merge, ONE launch.
// Every rank owns a window in HBM that all ranks of the box have mapped (CUDA IPC over NVLink/NVSwitch):
//   rec[2][world][kmax][2] u64    rec[s][r][i] = the i-th key of rank r for a step of parity s, as two words
//                                 (sequence number << 32 | upper half) and (sequence number << 32 | lower half)
// The CTA stores this rank's k keys straight into every peer's window -- 16 bytes per key and peer -- and every
// word carries the step's sequence number, so data and "it has arrived" travel in the same 8-byte store: no
// fence, no separate flag, one NVLink traversal (the low-latency protocol of collective libraries).  Each
// thread then polls one record of its OWN window until both words show this step's sequence number, and the
// world x k keys are merged with a rank sort.  No NCCL launch, no second merge launch, nothing through the host.
// Two parity slots make reuse safe without relying on timing: a rank cannot finish step s+1 before every peer
// has sent its step s+1 list, which a peer only does after it has consumed step s.
// The wait is bounded (~2 s of %globaltimer): a missing peer yields *res_count = kPeerTimeout, not a hung GPU.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelectThreads)
peer_exchange_merge_kernel(PeerExchange px, uint32_t seq, uint32_t k, int tie_higher, uint64_t *res_keys,
                           uint32_t *res_idx, float *res_val, uint32_t *res_count) {
    __shared__ uint64_t keys[kSelectSortCap];
    __shared__ uint32_t s_timeout, s_present;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { s_timeout = 0; s_present = 0; }
    pdl_wait();   // this rank's list comes from the select kernel right before
    const uint32_t mine_n = *res_count;
    if (tid < k) peer_push_key(px, seq, tid, (tid < mine_n) ? res_keys[tid] : 0ull);
    __syncthreads();   // s_timeout / s_present are initialised; every read of res_keys is done
    peer_poll_and_merge(px, seq, k, tie_higher, keys, &s_timeout, &s_present, res_keys, res_idx, res_val, res_count);
}

}  // namespace tks
