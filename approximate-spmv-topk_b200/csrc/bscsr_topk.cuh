// bscsr_topk.cuh -- FPGA-semantics Top-K SpMV on BS-CSR packets, for sm_100a.
//
// Reproduces bit for bit what the reference's HLS kernel computes for every partition
// (src/fpga/src/ip/spmv/spmv_bscsr_top_k_multicore.hpp:104-149, 168-220, 246-326, 331-409 and
// .cpp:112-185): W-bit unsigned fixed point with truncation and wrap, at most
// LIMITED_FINISHED_ROWS (LFR) row segments per packet, the row counter that merely COUNTS rows
// (including its drift when a packet holds more than LFR segments), the last row of a partition
// never offered, and LFR independent K-entry replace-min lists per partition with the argmin tie
// rule "highest slot among equal minima" (and the reference's argmin_4 typo).
//
// The reference streams each partition strictly in order (II=1 pipeline); 32 sequential streams
// cannot fill 148 SMs, so the stream is cut into chunks that are processed concurrently:
//
//   bscsr_stream_kernel   one warp per chunk, one thread per 64-byte packet per iteration:
//        decode, fixed-point products against the query in shared memory, <= LFR segment sums,
//        then two warp scans rebuild what the sequential kernel carries from packet to packet
//        (row counter; partial sum of the row that straddles packets).  The carry entering a
//        chunk is recomputed from the few packets before it (look-back precomputed at upload);
//        the row counter entering a chunk depends only on the matrix and is tabulated at upload.
//        Every candidate (value, row) that is >= the chunk's own running K-th largest is LOGGED in
//        stream order.  Any candidate the sequential kernel would have accepted is in the log,
//        because the sequential threshold (K-th largest of ALL earlier candidates) can only be
//        higher than the chunk-local one; rejected candidates never change the lists.
//   bscsr_replay_kernel   one CTA per (partition, lane): drops log entries below the K-th largest
//        of all EARLIER chunks (same argument), then replays the few hundred survivors through the
//        literal replace-min state machine.  Output: the reference's result words.
#pragma once

#include "common.cuh"

namespace tks {

constexpr uint32_t kBsThreads = 256;          // 8 warps per CTA in the stream kernel
constexpr uint32_t kBsMaxKp = 32;             // local K (types.hpp K) supported: 1..32
constexpr uint32_t kBsMaxLfr = 4;             // LFR values instantiated: 1..4 (see bscsr_api.cu)
constexpr uint32_t kReplayThreads = 256;
constexpr uint32_t kReplayTile = 512;         // chunks handled per tile in the replay kernel
constexpr uint32_t kReplaySurvivors = 2048;   // survivors buffered between sequential replays

struct BscsrDevice {
    const uint8_t *packets;          // all partitions back to back, 64 bytes per packet
    const uint32_t *chunk_first;     // global index of the chunk's first packet
    const uint32_t *chunk_count;     // packets in the chunk (<= chunk_cap)
    const uint32_t *chunk_local0;    // index of that packet inside its partition
    const uint32_t *chunk_row_in;    // the kernel's row counter before the chunk (upload-time table)
    const uint32_t *chunk_lookback;  // packets before the chunk needed to rebuild the carried partial sum
    uint32_t n_chunks;
    uint32_t chunk_cap;
};

struct BscsrLogs {
    uint32_t *val;     // [n_chunks][LFR][chunk_cap]
    uint32_t *row;     // [n_chunks][LFR][chunk_cap]
    uint32_t *cnt;     // [n_chunks][LFR]
    uint32_t *top;     // [n_chunks][LFR][32]  chunk-local K largest values, descending
    uint32_t *p0;      // [n_chunks][LFR]  1 when packet 0 of the PARTITION offered a candidate to lane j
};

template <int W>
struct BsFmt {
    static constexpr int B = 511 / (W + 14);          // types.hpp:71-72
    static constexpr int XOFF = 0, YOFF = 4 * B, VOFF = 14 * B;
    static constexpr uint32_t M = (W == 32) ? 0xFFFFFFFFu : ((1u << (W & 31)) - 1u);
};

// bits [lo, lo+width) of a 512-bit little-endian word held in 16 registers (compile-time position)
template <int LO, int WIDTH>
__device__ __forceinline__ uint32_t bs_field(const uint32_t (&w)[16]) {
    constexpr int q = LO / 32, s = LO % 32;
    constexpr uint32_t mask = (WIDTH == 32) ? 0xFFFFFFFFu : ((1u << (WIDTH & 31)) - 1u);
    if constexpr (s + WIDTH <= 32) {
        return (w[q] >> s) & mask;
    } else {
        return __funnelshift_r(w[q], w[q + 1], s) & mask;
    }
}

// sorted-descending insert of v into a K-entry list spread over lanes 0..Kp-1
__device__ __forceinline__ uint32_t lane_list_insert(uint32_t top, uint32_t v, uint32_t Kp) {
    const unsigned lane = lane_id();
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, top, 1);
    if (lane < Kp && top < v) top = (lane == 0 || up >= v) ? v : up;
    return top;
}

template <int W, int LFR>
__global__ void __launch_bounds__(kBsThreads)
bscsr_stream_kernel(BscsrDevice m, const uint32_t *__restrict__ xq, uint32_t Kp, BscsrLogs logs,
                    uint32_t *chunk_counter) {
    using F = BsFmt<W>;
    constexpr int B = F::B;
    constexpr uint32_t M = F::M;
    __shared__ uint32_t xs[1024];   // query, pre-shifted (see bscsr_api.cu); columns >= cols hold 0
    __shared__ uint32_t ptab[16 * kBsThreads];   // [prefix length 0..15][thread]: running sums of the products
    for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = xq[i];
    ptab[threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = lane_id();
    const uint8_t *xsb = reinterpret_cast<const uint8_t *>(xs);

    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(chunk_counter, 1u);
        c = __shfl_sync(0xFFFFFFFFu, c, 0);
        if (c >= m.n_chunks) break;
        const uint32_t first = m.chunk_first[c], count = m.chunk_count[c], local0 = m.chunk_local0[c];
        const uint32_t look = m.chunk_lookback[c];
        uint32_t row_base = m.chunk_row_in[c];   // last_row_of_packet before the next packet (hpp:260)
        uint32_t carry = 0;                      // last_row_of_packet_output (hpp:261)
        uint32_t theta[LFR], top[LFR], lcnt[LFR];
#pragma unroll
        for (int j = 0; j < LFR; j++) { theta[j] = 0; top[j] = 0; lcnt[j] = 0; }

        const uint32_t begin = first - look, end = first + count;
        for (uint32_t base = begin; base < end; base += 32) {
            const uint32_t g = base + lane;
            const bool active = g < end;
            const bool emitting = active && g >= first;
            uint32_t w[16];
            if (active) {
                const uint8_t *p = m.packets + (size_t)g * 64u;
                asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                             : "l"(p));
                asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                             : "l"(p + 32));
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) w[i] = 0;
            }
            // ---- loop 1 + loop 2 (hpp:168-220, 104-149): decode, products, segment sums ----
            uint32_t x[LFR];
            {
                // cumulative segment ends x[0..LFR-1]: 4-bit fields at bit 4*s
                const uint32_t x01 = w[0];
#pragma unroll
                for (int s = 0; s < LFR; s++) x[s] = (x01 >> (4 * s)) & 0xFu;   // LFR <= 8 fits the first word
            }
            // prefix sums of the products go to a per-thread column of shared memory, so that the LFR
            // "sum of the first x[s] products" are 4 conflict-free loads instead of 4 x B predicated adds
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < B; j++) {
                uint32_t pw;
                const uint32_t yoff = [&]() {
                    // column * 4 = byte offset into the query table
                    constexpr int lo = F::YOFF;
                    const int pos = lo + 10 * j;
                    const int q = pos / 32, s = pos % 32;
                    uint32_t f = (s + 10 <= 32) ? (w[q] >> s) : __funnelshift_r(w[q], w[q + 1], s);
                    return (f & 0x3FFu) << 2;
                }();
                const uint32_t xv = *reinterpret_cast<const uint32_t *>(xsb + yoff);
                const uint32_t v = [&]() {
                    const int pos = F::VOFF + W * j;
                    const int q = pos / 32, s = pos % 32;
                    uint32_t f = (s + W <= 32) ? (w[q] >> s) : __funnelshift_r(w[q], w[q + 1], s);
                    return f & M;
                }();
                if constexpr (W == 32) {
                    // ufixed<32,1> * ufixed<32,1> -> drop 31 fraction bits, wrap to 32 (hpp:121-126)
                    pw = (uint32_t)(((uint64_t)v * (uint64_t)xv) >> 31);
                } else {
                    // xs holds xq << 1 and v is top-aligned: umulhi gives (v * xq) >> (W-1) exactly
                    pw = __umulhi(v << (32 - W), xv);
                }
                acc += pw;
                ptab[(j + 1) * kBsThreads + threadIdx.x] = acc;
            }
            uint32_t Lsum[LFR];
#pragma unroll
            for (int s = 0; s < LFR; s++) Lsum[s] = ptab[x[s] * kBsThreads + threadIdx.x];   // x is non-decreasing (checked at upload)
            uint32_t agg[LFR];
            uint32_t n = 0;
#pragma unroll
            for (int s = 0; s < LFR; s++) {
                const uint32_t prev_end = s ? x[s - 1] : 0u;
                agg[s] = (Lsum[s] - (s ? Lsum[s - 1] : 0u)) & M;
                n += (x[s] != prev_end);
            }
            // ---- loop 3 (hpp:246-326): what is carried from packet to packet ----
            const uint32_t local_idx = local0 + (g - first);   // wraps correctly for look-back packets
            const uint32_t nw = (active && local_idx != 0) ? (w[15] >> 31) : 0u;
            // last_out recurrence  last_out_i = a_i + (k_i ? last_out_{i-1} : 0):
            //   n == 0: al[0]            -> (0, new)
            //   n == 1: al[1]            -> (agg0, !new)
            //   n >= 2: al[n] = agg[n-1] -> (agg[n-1], false)
            uint32_t a = 0;
            bool keep = true;   // identity for inactive lanes
            if (active) {
                if (n == 0) { a = 0; keep = nw != 0; }
                else if (n == 1) { a = agg[0]; keep = nw == 0; }
                else {
                    a = agg[LFR - 1];
#pragma unroll
                    for (int s = LFR - 2; s >= 1; s--) if (n == (uint32_t)(s + 1)) a = agg[s];
                    keep = false;
                }
            }
            uint32_t delta = emitting ? (n + nw - 1u) : 0u;    // finished_rows_num (hpp:281), u32 wrap
            // inclusive scans over the 32 packets of this iteration
            uint32_t sa = a;
            bool sk = keep;
            uint32_t sd = delta;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t ua = __shfl_up_sync(0xFFFFFFFFu, sa, d);
                const int uk = __shfl_up_sync(0xFFFFFFFFu, (int)sk, d);
                const uint32_t ud = __shfl_up_sync(0xFFFFFFFFu, sd, d);
                if ((int)lane >= d) {
                    if (sk) sa += ua;
                    sk = sk && (uk != 0);
                    sd += ud;
                }
            }
            // exclusive carry for this packet = inclusive value of the previous lane, chained to the warp carry
            uint32_t pa = __shfl_up_sync(0xFFFFFFFFu, sa, 1);
            int pk = __shfl_up_sync(0xFFFFFFFFu, (int)sk, 1);
            if (lane == 0) { pa = 0; pk = 1; }
            const uint32_t prev = (pa + (pk ? carry : 0u)) & M;          // last_row_of_packet_output seen by this packet
            const uint32_t start_row = row_base + (sd - delta) + nw;      // hpp:282
            // warp carries for the next iteration
            const uint32_t la = __shfl_sync(0xFFFFFFFFu, sa, 31);
            const int lk = __shfl_sync(0xFFFFFFFFu, (int)sk, 31);
            carry = (la + (lk ? carry : 0u)) & M;
            row_base += __shfl_sync(0xFFFFFFFFu, sd, 31);

            // ---- loop 4 (hpp:331-389): candidates of the LFR lanes ----
#pragma unroll
            for (int j = 0; j < LFR; j++) {
                uint32_t val;
                bool fin;
                if (j == 0) {
                    val = prev;            // al[0] = last_out when the packet starts a new row
                    fin = nw != 0;
                } else {
                    val = agg[j - 1];
                    if (j == 1 && nw == 0) val = (val + prev) & M;
                    fin = (x[j - 1] != (j > 1 ? x[j - 2] : 0u)) && (n != (uint32_t)j);
                }
                const uint32_t row = start_row + (uint32_t)j - 1u;
                if (active && local_idx == 0) logs.p0[(size_t)c * LFR + j] = fin ? 1u : 0u;
                const bool pass = emitting && fin && (val >= theta[j]);
                const unsigned pm = __ballot_sync(0xFFFFFFFFu, pass);
                if (pm) {
                    const size_t lbase = ((size_t)c * LFR + j) * m.chunk_cap;
                    if (pass) {
                        const uint32_t pos = lcnt[j] + __popc(pm & lanemask_lt());
                        logs.val[lbase + pos] = val;
                        logs.row[lbase + pos] = row;
                    }
                    lcnt[j] += __popc(pm);
                    unsigned rest = pm;
                    while (rest) {
                        const int src = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const uint32_t v = __shfl_sync(0xFFFFFFFFu, val, src);
                        top[j] = lane_list_insert(top[j], v, Kp);
                    }
                    theta[j] = __shfl_sync(0xFFFFFFFFu, top[j], (int)Kp - 1);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < LFR; j++) {
            if (lane == 0) logs.cnt[(size_t)c * LFR + j] = lcnt[j];
            logs.top[((size_t)c * LFR + j) * 32u + lane] = (lane < Kp) ? top[j] : 0u;
        }
    }
}

// kernel vec load (.cpp:127-137) for a query already in HBM: W-bit truncation of the raw 32-bit words,
// pre-shifted for the umulhi product; columns >= cols read 0 like the zero-initialised URAM copies.
template <int W>
__global__ void bscsr_query_kernel(const uint32_t *__restrict__ vec32, uint32_t cols, uint32_t *__restrict__ xq) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 1024) return;
    const uint32_t q = (c < cols) ? (vec32[c] >> (32 - W)) : 0u;
    xq[c] = (W == 32) ? q : (q << 1);
}

// argmin with MIN(res,a,b) = res[a] < res[b] ? a : b (hpp:28): highest slot among equal minima;
// K == 4 reproduces `MIN(res, 2, 2)` (hpp:45): slot 3 is never the minimum.
__device__ __forceinline__ uint32_t bs_argmin(const uint32_t *v, uint32_t Kp) {
    if (Kp == 4) {
        const uint32_t m0 = (v[0] < v[1]) ? 0u : 1u;
        return (v[m0] < v[2]) ? m0 : 2u;
    }
    uint32_t best = 0;
    for (uint32_t i = 1; i < Kp; i++) best = (v[best] < v[i]) ? best : i;
    return best;
}

template <int W>
__global__ void __launch_bounds__(kReplayThreads)
bscsr_replay_kernel(BscsrLogs logs, const uint32_t *__restrict__ part_chunk_begin, uint32_t LFR, uint32_t Kp,
                    uint32_t chunk_cap, uint32_t *res_idx_words, uint32_t *res_val_words, uint32_t *chunk_counter_reset) {
    const uint32_t p = blockIdx.x / LFR, j = blockIdx.x % LFR;
    const uint32_t cb = part_chunk_begin[p], ce = part_chunk_begin[p + 1];
    const uint32_t tid = threadIdx.x;
    const unsigned lane = lane_id();
    __shared__ uint32_t s_cnt[kReplayTile], s_off[kReplayTile + 1], s_thr[kReplayTile];
    __shared__ uint32_t s_top[kReplayTile * 8];          // staged chunk tops, 8 at a time (see below)
    __shared__ uint32_t s_sv[kReplaySurvivors], s_sr[kReplaySurvivors];
    __shared__ uint32_t s_n, s_wsum[kReplayThreads / 32];
    __shared__ uint32_t Lval[kBsMaxKp], Lidx[kBsMaxKp];
    __shared__ uint32_t s_worst_idx, s_worst_val, s_started;
    if (tid < kBsMaxKp) { Lval[tid] = 0; Lidx[tid] = 0; }
    if (tid == 0) { s_worst_idx = 0; s_worst_val = 0; s_started = 0; s_n = 0; }
    if (blockIdx.x == 0 && tid == 0 && chunk_counter_reset) *chunk_counter_reset = 0;
    uint32_t rtop = 0;   // warp 0: running K largest values of all chunks seen so far (lanes 0..Kp-1)
    const bool first_from_packet0 = logs.p0[(size_t)cb * LFR + j] != 0;
    __syncthreads();

    auto replay = [&]() {
        // sequential, literal (hpp:366-389): only thread 0
        if (tid == 0) {
            uint32_t wi = s_worst_idx, wv = s_worst_val, started = s_started;
            const uint32_t n = s_n;
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t v = s_sv[i], r = s_sr[i];
                if (!started) {
                    // the argmin is recomputed after EVERY packet (hpp:376-388): unless the first candidate comes
                    // from packet 0 of the partition, the all-zero list has already moved the worst slot
                    if (!first_from_packet0) { wi = bs_argmin(Lval, Kp); wv = Lval[wi]; }
                    started = 1;
                }
                if (v >= wv) {
                    Lidx[wi] = r;
                    Lval[wi] = v;
                    wi = bs_argmin(Lval, Kp);
                    wv = Lval[wi];
                }
            }
            s_worst_idx = wi; s_worst_val = wv; s_started = started; s_n = 0;
        }
        __syncthreads();
    };

    for (uint32_t t0 = cb; t0 < ce; t0 += kReplayTile) {
        const uint32_t nt = (ce - t0 < kReplayTile) ? (ce - t0) : kReplayTile;
        for (uint32_t i = tid; i < nt; i += blockDim.x) s_cnt[i] = logs.cnt[(size_t)(t0 + i) * LFR + j];
        __syncthreads();
        if (tid == 0) {
            uint32_t acc = 0;
            for (uint32_t i = 0; i < nt; i++) { s_off[i] = acc; acc += s_cnt[i]; }
            s_off[nt] = acc;
        }
        // thresholds: K-th largest value among the tops of all EARLIER chunks (warp 0, chunks in order)
        for (uint32_t sub = 0; sub < nt; sub += 128) {   // tops staged 128 chunks at a time: 128 x 32 words = s_top? no: 8 words used
            __syncthreads();
            const uint32_t ns = (nt - sub < 128) ? (nt - sub) : 128;
            // stage Kp (<= 32) tops of ns chunks; s_top holds ns * 32 words -> reuse as [128][32]
            for (uint32_t i = tid; i < ns * 32u; i += blockDim.x)
                s_top[i] = logs.top[((size_t)(t0 + sub + i / 32u) * LFR + j) * 32u + (i % 32u)];
            __syncthreads();
            if (tid < 32) {
                for (uint32_t cidx = 0; cidx < ns; cidx++) {
                    const uint32_t thr = __shfl_sync(0xFFFFFFFFu, rtop, (int)Kp - 1);
                    if (lane == 0) s_thr[sub + cidx] = thr;
                    const uint32_t v = s_top[cidx * 32u + lane];
                    unsigned rest = __ballot_sync(0xFFFFFFFFu, lane < Kp && v > thr);
                    while (rest) {
                        const int src = __ffs(rest) - 1;
                        rest &= rest - 1;
                        rtop = lane_list_insert(rtop, __shfl_sync(0xFFFFFFFFu, v, src), Kp);
                    }
                }
            }
        }
        __syncthreads();
        // order-preserving filter of the tile's concatenated logs
        const uint32_t total = s_off[nt];
        for (uint32_t base = 0; base < total; base += blockDim.x) {
            const uint32_t f = base + tid;
            bool keep = false;
            uint32_t v = 0, r = 0;
            if (f < total) {
                uint32_t lo = 0, hi = nt;   // chunk with s_off[chunk] <= f < s_off[chunk+1]
                while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_off[mid] <= f) lo = mid; else hi = mid; }
                const size_t e = ((size_t)(t0 + lo) * LFR + j) * chunk_cap + (f - s_off[lo]);
                v = logs.val[e];
                r = logs.row[e];
                keep = v >= s_thr[lo];
            }
            const unsigned km = __ballot_sync(0xFFFFFFFFu, keep);
            if (lane == 0) s_wsum[tid / 32] = __popc(km);
            __syncthreads();
            uint32_t before = 0, all = 0;
            for (uint32_t wv = 0; wv < blockDim.x / 32; wv++) { if (wv < tid / 32) before += s_wsum[wv]; all += s_wsum[wv]; }
            if (s_n + all > kReplaySurvivors) {   // uniform: flush what is buffered first
                __syncthreads();
                replay();
            }
            if (keep) {
                const uint32_t pos = s_n + before + __popc(km & lanemask_lt());
                s_sv[pos] = v;
                s_sr[pos] = r;
            }
            __syncthreads();
            if (tid == 0) s_n += all;
            __syncthreads();
        }
        replay();
    }
    // write-back (.cpp:151-185): word t, position j = list j slot t; values widened to ufixed<32,1>
    if (tid < Kp) {
        const size_t o = ((size_t)p * Kp + tid) * 16u + j;
        res_idx_words[o] = Lidx[tid];
        res_val_words[o] = (W == 32) ? Lval[tid] : (Lval[tid] << (32 - W));
    }
}

}  // namespace tks
