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
//   bscsr_sample_kernel   reduces the first packets of every partition (in 64-packet pieces, one per
//        warp) and leaves, per (partition, lane), the K-th largest candidate seen there.  Those are
//        real candidates that precede every later chunk of the partition in stream order, so the
//        value is a lower bound on the sequential kernel's threshold for all those chunks.
//   bscsr_stream_kernel   one warp per chunk, one thread per 64-byte packet per iteration:
//        decode, fixed-point products against the query in shared memory, <= LFR segment sums,
//        then two warp scans rebuild what the sequential kernel carries from packet to packet
//        (row counter; partial sum of the row that straddles packets).  The carry entering a
//        chunk is recomputed from the few packets before it (look-back precomputed at upload);
//        the row counter entering a chunk depends only on the matrix and is tabulated at upload.
//        Every candidate (value, row) that is >= max(seed, the chunk's own running K-th largest)
//        is LOGGED in stream order.  Any candidate the sequential kernel would have accepted is in
//        the log, because the sequential threshold (K-th largest of ALL earlier candidates) can only
//        be higher; rejected candidates never change the lists.
//   bscsr_replay_kernel   one CTA per (partition, lane): gathers the chunk logs in order and replays
//        them through the literal replace-min state machine.  Output: the reference's result words.
#pragma once

#include "common.cuh"

namespace tks {

constexpr uint32_t kBsThreads = 256;          // 8 warps per CTA in the sample kernel
constexpr int kBsDefaultXrep = 32;            // stream kernel: one query copy per shared-memory bank
constexpr int kBsDefaultThreads = 768;        // stream kernel: one CTA of 24 warps per SM (128 KB query + 48 KB ptab)
constexpr uint32_t kBsMaxKp = 32;             // local K (types.hpp K) supported: 1..32
constexpr uint32_t kBsMaxLfr = 4;             // LFR values instantiated: 1..4 (see bscsr_api.cu)
constexpr uint32_t kBsSamplePackets = 2048;   // prefix of every partition reduced by the sample kernel (at most)
constexpr uint32_t kBsSamplePiece = 64;       // packets per sample warp, look-back included (two warp iterations)
constexpr uint32_t kBsMaxPieces = 40;         // sample pieces per partition (bounds the merge's register array)
constexpr uint32_t kReplayThreads = 1024;      // one tile covers the ~800 chunks of a cfg3 partition
constexpr uint32_t kReplaySurvivors = 8192;   // log entries buffered between sequential replays (dynamic smem)
constexpr uint32_t kReplayDynSmem = kReplaySurvivors * 8u;
constexpr uint32_t kReplayCpt = 4;             // chunks per replay thread and tile

struct BscsrChunks {
    const uint32_t *first;      // global index of the chunk's first packet
    const uint32_t *count;      // packets in the chunk
    const uint32_t *local0;     // index of that packet inside its partition
    const uint32_t *row_in;     // the kernel's row counter before the chunk (upload-time table)
    const uint32_t *lookback;   // packets before the chunk needed to rebuild the carried partial sum
    const uint32_t *part;       // partition of the chunk
    uint32_t n;
    uint32_t cap;               // max packets per chunk = log capacity per (chunk, lane)
};

struct BscsrLogs {
    uint32_t *val;     // [n_chunks][LFR][cap]
    uint32_t *row;     // [n_chunks][LFR][cap]
    uint32_t *cnt;     // [n_chunks][LFR]
    uint32_t *p0;      // [n_chunks][LFR]  1 when packet 0 of the PARTITION offered a candidate to lane j
};

template <int W>
struct BsFmt {
    static constexpr int B = 511 / (W + 14);          // types.hpp:71-72
    static constexpr int YOFF = 4 * B, VOFF = 14 * B;
    static constexpr uint32_t M = (W == 32) ? 0xFFFFFFFFu : ((1u << (W & 31)) - 1u);
};

// sorted-descending insert of v into a K-entry list spread over lanes 0..Kp-1
__device__ __forceinline__ uint32_t lane_list_insert(uint32_t top, uint32_t v, uint32_t Kp) {
    const unsigned lane = lane_id();
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, top, 1);
    if (lane < Kp && top < v) top = (lane == 0 || up >= v) ? v : up;
    return top;
}

__device__ __forceinline__ void bs_load_packet(const uint8_t *p, uint32_t (&w)[16]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                 : "l"(p + 32));
}

// Sinks of the packet loop.  LogSink keeps the stream-ordered log of one chunk; TopSink (sample) only
// keeps the K largest values per lane.
template <int LFR>
struct BsLogSink {
    uint32_t *val, *row;   // this chunk's log, [LFR][cap]
    uint32_t cap;
    uint32_t lcnt[LFR];
    __device__ __forceinline__ void put(int j, bool pass, unsigned pm, uint32_t v, uint32_t r) {
        if (pass) {
            const uint32_t pos = lcnt[j] + __popc(pm & lanemask_lt());
            val[(size_t)j * cap + pos] = v;
            row[(size_t)j * cap + pos] = r;
        }
        lcnt[j] += __popc(pm);
    }
};
template <int LFR>
struct BsTopSink {
    __device__ __forceinline__ void put(int, bool, unsigned, uint32_t, uint32_t) {}
};

// Packets [begin, end) of one partition, 32 per iteration; packets before `first` only rebuild the carry.
// theta/top: per lane-list running K-th largest and the K largest values (lanes 0..Kp-1), updated in place.
// XREP: copies of the query in shared memory (word col * XREP + (lane % XREP)): with 32 copies every lane
// gathers from its own bank (no conflicts), with 16 two lanes share a bank pair, with 1 the gather is the
// plain 1024-word table.  THREADS: CTA size = stride of the per-thread prefix-sum columns in ptab.
// PREFETCH: keep the next iteration's packet in a second register set (16 more registers per thread).
template <int W, int LFR, int XREP, int THREADS, bool PREFETCH, typename Sink>
__device__ __forceinline__ void bs_process(const uint8_t *__restrict__ packets, uint32_t begin, uint32_t first,
                                           uint32_t end, uint32_t local0, uint32_t row_base, uint32_t Kp,
                                           const uint8_t *xsb, uint32_t *ptab, uint32_t (&theta)[LFR],
                                           uint32_t (&top)[LFR], Sink &sink, uint32_t *p0flags) {
    using F = BsFmt<W>;
    constexpr int B = F::B;
    constexpr uint32_t M = F::M;
    constexpr int XS = (XREP == 32) ? 7 : (XREP == 16) ? 6 : (XREP == 8) ? 5 : (XREP == 4) ? 4 : (XREP == 2) ? 3 : 2;
    static_assert((1 << XS) == 4 * XREP, "XREP must be a power of two <= 32");
    constexpr uint32_t YMASK = 0x3FFu << XS;
    const unsigned lane = lane_id();
    xsb += (lane & (uint32_t)(XREP - 1)) * 4u;   // this lane's copy
    uint32_t carry = 0;   // last_row_of_packet_output (hpp:261)
    uint32_t wa[16], wb[PREFETCH ? 16 : 1];
    if (PREFETCH) {
        const uint32_t g0 = begin + lane;
        if (g0 < end) bs_load_packet(packets + (size_t)g0 * 64u, wa);
        else {
#pragma unroll
            for (int i = 0; i < 16; i++) wa[i] = 0;
        }
    }
    // one iteration = 32 consecutive packets, one per lane; `w` holds this iteration's packet, `wn` receives the
    // next one (software prefetch).  The caller alternates the two register sets instead of copying them.
    auto iteration = [&](uint32_t base, uint32_t (&w)[16], uint32_t (&wn)[PREFETCH ? 16 : 1]) {
        const uint32_t g = base + lane;
        const bool active = g < end;
        const bool emitting = active && g >= first;
        if constexpr (PREFETCH) {
            const uint32_t gn = g + 32;
            if (gn < end) bs_load_packet(packets + (size_t)gn * 64u, wn);
            else {
#pragma unroll
                for (int i = 0; i < 16; i++) wn[i] = 0;
            }
        } else {
            if (active) bs_load_packet(packets + (size_t)g * 64u, w);
            else {
#pragma unroll
                for (int i = 0; i < 16; i++) w[i] = 0;
            }
        }
        // ---- loop 1 + loop 2 (hpp:168-220, 104-149): decode, products, segment sums ----
        uint32_t x[LFR];
#pragma unroll
        for (int s = 0; s < LFR; s++) x[s] = (w[0] >> (4 * s)) & 0xFu;   // cumulative segment ends (LFR <= 8)
        // prefix sums of the products go to a per-thread column of shared memory, so that the LFR
        // "sum of the first x[s] products" are LFR conflict-free loads instead of LFR x B predicated adds
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < B; j++) {
            // column * 4 (byte offset into the query table) and the value: one shift + one mask each
            const int yp = F::YOFF + 10 * j, yq = yp / 32, ysh = yp % 32;
            uint32_t yraw;
            if (ysh + 10 <= 32) yraw = (ysh >= XS) ? (w[yq] >> (ysh - XS)) : (w[yq] << (XS - ysh));
            else yraw = __funnelshift_r(w[yq], w[yq + 1], ysh - XS);
            const uint32_t xv = *reinterpret_cast<const uint32_t *>(xsb + (yraw & YMASK));
            const int vp = F::VOFF + W * j, vq = vp / 32, vsh = vp % 32;
            uint32_t pw;
            if constexpr (W == 32) {
                // ufixed<32,1> * ufixed<32,1> -> drop 31 fraction bits, wrap to 32 (hpp:121-126)
                const uint32_t v = (vsh == 0) ? w[vq] : __funnelshift_r(w[vq], w[vq + 1], vsh);
                pw = (uint32_t)(((uint64_t)v * (uint64_t)xv) >> 31);
            } else {
                // the query table holds xq << 1 and v is moved to the top of the word:
                // umulhi gives (v * xq) >> (W-1) exactly
                constexpr uint32_t topmask = ~((1u << (32 - W)) - 1u);
                uint32_t vraw;
                if (vsh + W <= 32) vraw = w[vq] << (32 - W - vsh);
                else vraw = __funnelshift_r(w[vq], w[vq + 1], vsh - (32 - W));
                pw = __umulhi(vraw & topmask, xv);
            }
            acc += pw;
            ptab[(j + 1) * THREADS + threadIdx.x] = acc;
        }
        uint32_t agg[LFR];
        uint32_t n = 0;
        {
            uint32_t prevL = 0, prev_end = 0;
#pragma unroll
            for (int s = 0; s < LFR; s++) {
                const uint32_t L = ptab[x[s] * THREADS + threadIdx.x];   // x is non-decreasing (checked at upload)
                agg[s] = (L - prevL) & M;
                n += (x[s] != prev_end);
                prevL = L;
                prev_end = x[s];
            }
        }
        // ---- loop 3 (hpp:246-326): what is carried from packet to packet ----
        const uint32_t local_idx = local0 + (g - first);   // wraps correctly for look-back packets
        const uint32_t nw = (active && local_idx != 0) ? (w[15] >> 31) : 0u;
        // last_out recurrence  last_out_i = a_i + (keep_i ? last_out_{i-1} : 0):
        //   n == 0: al[0]            -> (0, new)
        //   n == 1: al[1]            -> (agg0, !new)
        //   n >= 2: al[n] = agg[n-1] -> (agg[n-1], false)
        uint32_t a = 0;
        uint32_t keep = 1;   // identity for inactive lanes
        if (active) {
            if (n == 0) { a = 0; keep = nw; }
            else if (n == 1) { a = agg[0]; keep = nw ^ 1u; }
            else {
                a = agg[LFR - 1];
#pragma unroll
                for (int s = LFR - 2; s >= 1; s--) if (n == (uint32_t)(s + 1)) a = agg[s];
                keep = 0;
            }
        }
        const uint32_t delta = emitting ? (n + nw - 1u) : 0u;    // finished_rows_num (hpp:281)
        // inclusive scans over the 32 packets of this iteration; the row-counter increments (< 2^16 per
        // iteration) and the keep flag share one register: bits 0..15 sum of delta, bit 16 keep
        uint32_t sa = a;
        uint32_t sdk = (delta & 0xFFFFu) | (keep << 16);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ua = __shfl_up_sync(0xFFFFFFFFu, sa, d);
            const uint32_t udk = __shfl_up_sync(0xFFFFFFFFu, sdk, d);
            if ((int)lane >= d) {
                if (sdk & 0x10000u) sa += ua;
                sdk = ((sdk + udk) & 0xFFFFu) | (sdk & udk & 0x10000u);
            }
        }
        // exclusive carry for this packet = inclusive value of the previous lane, chained to the warp carry
        uint32_t pa = __shfl_up_sync(0xFFFFFFFFu, sa, 1);
        uint32_t pdk = __shfl_up_sync(0xFFFFFFFFu, sdk, 1);
        if (lane == 0) { pa = 0; pdk = 0x10000u; }
        const uint32_t prev = (pa + ((pdk & 0x10000u) ? carry : 0u)) & M;   // last_row_of_packet_output seen by this packet
        const uint32_t start_row = row_base + (pdk & 0xFFFFu) + nw;          // hpp:282
        // warp carries for the next iteration
        const uint32_t la = __shfl_sync(0xFFFFFFFFu, sa, 31);
        const uint32_t ldk = __shfl_sync(0xFFFFFFFFu, sdk, 31);
        carry = (la + ((ldk & 0x10000u) ? carry : 0u)) & M;
        row_base += ldk & 0xFFFFu;

        // ---- loop 4 (hpp:331-389): candidates of the LFR lanes ----
        uint32_t val[LFR];
        bool pass[LFR], fin0[LFR];
        bool anyp = false;
#pragma unroll
        for (int j = 0; j < LFR; j++) {
            bool fin;
            if (j == 0) {
                val[0] = prev;            // al[0] = last_out when the packet starts a new row
                fin = nw != 0;
            } else {
                val[j] = agg[j - 1];
                if (j == 1 && nw == 0) val[j] = (val[j] + prev) & M;
                fin = (x[j - 1] != (j > 1 ? x[j - 2] : 0u)) && (n != (uint32_t)j);
            }
            fin0[j] = fin;
            pass[j] = emitting && fin && (val[j] >= theta[j]);
            anyp |= pass[j];
        }
        if (p0flags && base <= first && first < base + 32 && local0 == 0) {   // warp-uniform: packet 0 of the partition
            if (active && local_idx == 0) {
#pragma unroll
                for (int j = 0; j < LFR; j++) p0flags[j] = fin0[j] ? 1u : 0u;
            }
        }
        if (__any_sync(0xFFFFFFFFu, anyp)) {
#pragma unroll
            for (int j = 0; j < LFR; j++) {
                const unsigned pm = __ballot_sync(0xFFFFFFFFu, pass[j]);
                if (pm) {
                    sink.put(j, pass[j], pm, val[j], start_row + (uint32_t)j - 1u);
                    unsigned rest = pm;
                    while (rest) {
                        const int src = __ffs(rest) - 1;
                        rest &= rest - 1;
                        top[j] = lane_list_insert(top[j], __shfl_sync(0xFFFFFFFFu, val[j], src), Kp);
                    }
                    theta[j] = __shfl_sync(0xFFFFFFFFu, top[j], (int)Kp - 1);
                }
            }
        }
    };
    if constexpr (PREFETCH) {
        for (uint32_t base = begin; base < end; base += 64) {
            iteration(base, wa, wb);
            if (base + 32 < end) iteration(base + 32, wb, wa);
        }
    } else {
        for (uint32_t base = begin; base < end; base += 32) iteration(base, wa, wb);
    }
}

// --------------------------------------------------------------------------------------------------
// BSX: the device-side packet format for FIXED_WIDTH <= 22 (one non-zero per 32-bit word).
//
// Everything about a packet that does not depend on the query -- where its fields sit, how many row
// segments it closes, whether it starts a new row, the value of the kernel's row counter when it is
// reached -- is decided by the matrix alone, so tks_upload_bscsr re-encodes the reference's 512-bit words
// (fpga_utils.hpp:307-365) once, losslessly for what the kernel consumes, into 16 aligned words:
//     word j < B :  val_j << (32 - W)  |  col_j                    (W-bit value on top, 10-bit column at the bottom)
//     word 15    :  x[0..3] (4 bits each, cumulative segment ends)  |  meta << 16
//     meta       :  bit 0      nw   = (packet index != 0) ? xf : 0          (hpp:278)
//                   bits 1..3  n    = non-empty segments among the first LFR (hpp:131-142)
//                   bits 4..15 rel  = row counter before the packet, relative to its chunk, + nw (hpp:280-282)
// Per query this leaves: B x (mask, shift, gather, multiply-high, add, store), four prefix-sum look-ups, one
// segmented warp scan for the partial sum carried from packet to packet, and the threshold tests.
// Semantics are those of bs_process (same candidates, same order, same logs).
// --------------------------------------------------------------------------------------------------
template <int W>
struct BsxFmt {
    static_assert(W + 10 <= 32, "BSX needs value + column in one word");
    static constexpr int B = 511 / (W + 14);
    static constexpr uint32_t M = (1u << W) - 1u;
    static constexpr uint32_t TOPMASK = ~((1u << (32 - W)) - 1u);
    static constexpr int META = 15;
};

template <int W, int LFR, int XREP, int THREADS, bool PREFETCH, typename Sink>
__device__ __forceinline__ void bsx_process(const uint8_t *__restrict__ packets, uint32_t begin, uint32_t first,
                                            uint32_t end, uint32_t local0, uint32_t row_base, uint32_t Kp,
                                            const uint8_t *xsb, uint32_t *ptab, uint32_t (&theta)[LFR],
                                            uint32_t (&top)[LFR], Sink &sink, uint32_t *p0flags) {
    using F = BsxFmt<W>;
    constexpr int B = F::B;
    constexpr uint32_t M = F::M;
    constexpr int XS = (XREP == 32) ? 7 : (XREP == 16) ? 6 : (XREP == 8) ? 5 : (XREP == 4) ? 4 : (XREP == 2) ? 3 : 2;
    static_assert((1 << XS) == 4 * XREP, "XREP must be a power of two <= 32");
    static_assert(LFR <= 4, "word 15 carries x[0..3]");
    const unsigned lane = lane_id();
    // 32-bit shared-memory address of this lane's copy of the query: one multiply-add per gather address
    const uint32_t xs_addr = (uint32_t)__cvta_generic_to_shared(xsb) + (lane & (uint32_t)(XREP - 1)) * 4u;
    const unsigned lt = lanemask_lt();
    uint32_t carry = 0;   // last_row_of_packet_output (hpp:261) entering the iteration
    uint32_t wa[16], wb[PREFETCH ? 16 : 1];
    if (PREFETCH) {
        const uint32_t g0 = begin + lane;
        if (g0 < end) bs_load_packet(packets + (size_t)g0 * 64u, wa);
        else {
#pragma unroll
            for (int i = 0; i < 16; i++) wa[i] = 0;
        }
    }
    auto iteration = [&](uint32_t base, uint32_t (&w)[16], uint32_t (&wn)[PREFETCH ? 16 : 1]) {
        const uint32_t g = base + lane;
        const bool active = g < end;
        const bool emitting = active && g >= first;
        if constexpr (PREFETCH) {
            const uint32_t gn = g + 32;
            if (gn < end) bs_load_packet(packets + (size_t)gn * 64u, wn);
            else {
#pragma unroll
                for (int i = 0; i < 16; i++) wn[i] = 0;
            }
        } else {
            if (active) bs_load_packet(packets + (size_t)g * 64u, w);
            else {
#pragma unroll
                for (int i = 0; i < 16; i++) w[i] = 0;
            }
            // no second register set: pull the packets of the iteration after next into L2 instead
            if (g + 64 < end) {
                const uint8_t *pf = packets + (size_t)(g + 64) * 64u;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 32));
            }
        }
        // ---- loops 1 + 2 (hpp:168-220, 104-149): products and their running sums ----
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < B; j++) {
            uint32_t xv;
            asm("ld.shared.u32 %0, [%1];" : "=r"(xv) : "r"((w[j] & 0x3FFu) * (4u * XREP) + xs_addr));
            acc += __umulhi(w[j] & F::TOPMASK, xv);   // the table holds xq << 1: exactly (val * xq) >> (W-1)
            ptab[(j + 1) * THREADS + threadIdx.x] = acc;
        }
        const uint32_t mw = w[F::META];
        uint32_t x[LFR], agg[LFR];
        {
            uint32_t prevL = 0;
#pragma unroll
            for (int s = 0; s < LFR; s++) {
                x[s] = (mw >> (4 * s)) & 0xFu;
                const uint32_t L = ptab[x[s] * THREADS + threadIdx.x];
                agg[s] = (L - prevL) & M;
                prevL = L;
            }
        }
        const uint32_t nw = (mw >> 16) & 1u;
        const uint32_t n = (mw >> 17) & 7u;
        const uint32_t start_row = row_base + (mw >> 20);   // hpp:282, tabulated at upload
        // ---- loop 3 (hpp:246-326): the partial sum carried from packet to packet ----
        // last_out_i = a_i + (keep_i ? last_out_{i-1} : 0):  n == 0 -> (0, nw);  n == 1 -> (agg0, !nw);  n >= 2 -> (agg[n-1], 0)
        uint32_t a = 0;
#pragma unroll
        for (int s = 0; s < LFR; s++) if (n == (uint32_t)(s + 1)) a = agg[s];
        const bool keep = !active || (n < 2u && ((n ^ nw) & 1u));   // inactive lanes are the identity
        const unsigned kb = __ballot_sync(0xFFFFFFFFu, keep);
        // lanes below this one that belong to its run of `keep` packets (segmented inclusive scan by distance)
        const unsigned brk = ~kb & (lt | (1u << lane));              // non-keep lanes at or below this one
        const int dist = brk ? (int)lane - (31 - __clz((int)brk)) : (int)lane;
        uint32_t sa = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ua = __shfl_up_sync(0xFFFFFFFFu, sa, d);
            if (dist >= d) sa += ua;
        }
        // what this packet sees as last_row_of_packet_output: the inclusive value of the previous lane, chained to
        // the carry of the previous iteration when every lane below passes it through
        uint32_t pa = __shfl_up_sync(0xFFFFFFFFu, sa, 1);
        if (lane == 0) pa = 0;
        const uint32_t prev = (pa + (((~kb & lt) == 0u) ? carry : 0u)) & M;
        const uint32_t la = __shfl_sync(0xFFFFFFFFu, sa, 31);
        carry = (la + ((kb == 0xFFFFFFFFu) ? carry : 0u)) & M;

        // ---- loop 4 (hpp:331-389): candidates of the LFR lanes ----
        uint32_t val[LFR];
        bool pass[LFR], fin0[LFR];
        bool anyp = false;
#pragma unroll
        for (int j = 0; j < LFR; j++) {
            bool fin;
            if (j == 0) {
                val[0] = prev;
                fin = nw != 0;
            } else {
                val[j] = agg[j - 1];
                if (j == 1 && nw == 0) val[j] = (val[j] + prev) & M;
                fin = (x[j - 1] != (j > 1 ? x[j - 2] : 0u)) && (n != (uint32_t)j);
            }
            fin0[j] = fin;
            pass[j] = emitting && fin && (val[j] >= theta[j]);
            anyp |= pass[j];
        }
        if (p0flags && base <= first && first < base + 32 && local0 == 0) {   // warp-uniform: packet 0 of the partition
            if (g == first) {
#pragma unroll
                for (int j = 0; j < LFR; j++) p0flags[j] = fin0[j] ? 1u : 0u;
            }
        }
        if (__any_sync(0xFFFFFFFFu, anyp)) {
#pragma unroll
            for (int j = 0; j < LFR; j++) {
                const unsigned pm = __ballot_sync(0xFFFFFFFFu, pass[j]);
                if (pm) {
                    sink.put(j, pass[j], pm, val[j], start_row + (uint32_t)j - 1u);
                    unsigned rest = pm;
                    while (rest) {
                        const int src = __ffs(rest) - 1;
                        rest &= rest - 1;
                        top[j] = lane_list_insert(top[j], __shfl_sync(0xFFFFFFFFu, val[j], src), Kp);
                    }
                    theta[j] = __shfl_sync(0xFFFFFFFFu, top[j], (int)Kp - 1);
                }
            }
        }
    };
    if constexpr (PREFETCH) {
        for (uint32_t base = begin; base < end; base += 64) {
            iteration(base, wa, wb);
            if (base + 32 < end) iteration(base + 32, wb, wa);
        }
    } else {
        for (uint32_t base = begin; base < end; base += 32) iteration(base, wa, wb);
    }
}

// verbatim reference packets (BSX = false) or the re-encoded device format (BSX = true, FIXED_WIDTH <= 22)
template <int W, int LFR, int XREP, int THREADS, bool PREFETCH, bool BSX, typename Sink>
__device__ __forceinline__ void bs_run(const uint8_t *__restrict__ packets, uint32_t begin, uint32_t first, uint32_t end,
                                       uint32_t local0, uint32_t row_base, uint32_t Kp, const uint8_t *xsb,
                                       uint32_t *ptab, uint32_t (&theta)[LFR], uint32_t (&top)[LFR], Sink &sink,
                                       uint32_t *p0flags) {
    if constexpr (BSX)
        bsx_process<W, LFR, XREP, THREADS, PREFETCH>(packets, begin, first, end, local0, row_base, Kp, xsb, ptab, theta, top,
                                                     sink, p0flags);
    else
        bs_process<W, LFR, XREP, THREADS, PREFETCH>(packets, begin, first, end, local0, row_base, Kp, xsb, ptab, theta, top,
                                                    sink, p0flags);
}

struct BscsrSample {
    const uint32_t *first, *count, *local0, *lookback, *part;   // sample pieces (64 packets each)
    uint32_t n;
    const uint32_t *part_piece_begin;   // [P+1]
    uint32_t *piece_top;                // [n][LFR][32]
    uint32_t *ticket;                   // [P] pieces finished
    uint32_t *theta_seed;               // [P][LFR]
};

template <int W, int LFR, bool BSX>
__global__ void __launch_bounds__(kBsThreads)
bscsr_sample_kernel(const uint8_t *__restrict__ packets, BscsrSample sm, const uint32_t *__restrict__ xq, uint32_t Kp) {
    __shared__ uint32_t xs[1024];
    __shared__ uint32_t ptab[16 * kBsThreads];
    pdl_trigger();   // the stream kernel may start staging its query copies while the pieces are reduced
    for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = xq[i];
    ptab[threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = lane_id();
    const uint32_t piece = (blockIdx.x * blockDim.x + threadIdx.x) / 32u;
    if (piece >= sm.n) return;
    uint32_t theta[LFR], top[LFR];
#pragma unroll
    for (int j = 0; j < LFR; j++) { theta[j] = 0; top[j] = 0; }
    BsTopSink<LFR> sink;
    const uint32_t first = sm.first[piece];
    bs_run<W, LFR, 1, kBsThreads, false, BSX>(packets, first - sm.lookback[piece], first, first + sm.count[piece],
                                              sm.local0[piece], 0u, Kp, reinterpret_cast<const uint8_t *>(xs), ptab, theta,
                                              top, sink, nullptr);
#pragma unroll
    for (int j = 0; j < LFR; j++) sm.piece_top[((size_t)piece * LFR + j) * 32u + lane] = (lane < Kp) ? top[j] : 0u;
    // the last piece of the partition to finish merges the pieces' tops: K-th largest of the sampled prefix
    const uint32_t p = sm.part[piece];
    const uint32_t pb = sm.part_piece_begin[p], pe = sm.part_piece_begin[p + 1];
    __threadfence();
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(&sm.ticket[p], 1u);
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    if (t != pe - pb - 1) return;
    __threadfence();
    constexpr uint32_t kMaxPieces = kBsMaxPieces;
    if (Kp <= 8 && LFR <= 4) {
        // the LFR lists are merged side by side: lanes [8j, 8j+8) hold list j (slot = lane % 8), so one pass over the
        // pieces serves all lists; the pieces' values are fetched first (independent loads), then merged from registers
        const uint32_t jl = lane >> 3, t8 = lane & 7u, base8 = lane & ~7u;
        uint32_t vq[kMaxPieces];
#pragma unroll
        for (uint32_t q = 0; q < kMaxPieces; q++)
            vq[q] = (pb + q < pe && jl < (uint32_t)LFR && t8 < Kp)
                        ? __ldcg(&sm.piece_top[((size_t)(pb + q) * LFR + jl) * 32u + t8]) : 0u;
        uint32_t rtop = 0;
#pragma unroll
        for (uint32_t q = 0; q < kMaxPieces; q++) {
            uint32_t thr = __shfl_sync(0xFFFFFFFFu, rtop, (int)(base8 + Kp - 1));
            unsigned rest = __ballot_sync(0xFFFFFFFFu, vq[q] > thr);
            while (rest) {
                const uint32_t t = ((uint32_t)__ffs(rest) - 1u) & 7u;          // slot t of every list's piece
                rest &= ~(0x01010101u << t);
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, vq[q], (int)(base8 + t));
                thr = __shfl_sync(0xFFFFFFFFu, rtop, (int)(base8 + Kp - 1));
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, rtop, 1, 8);
                if (v > thr && t8 < Kp && rtop < v) rtop = (t8 == 0 || up >= v) ? v : up;
            }
        }
        const uint32_t seed = __shfl_sync(0xFFFFFFFFu, rtop, (int)(base8 + Kp - 1));
        if (t8 == 0 && jl < (uint32_t)LFR) sm.theta_seed[p * LFR + jl] = seed;
    } else {
#pragma unroll 1
        for (int j = 0; j < LFR; j++) {
            // all loads first (independent), then the sequential merge from registers
            uint32_t vq[kMaxPieces];
#pragma unroll
            for (uint32_t q = 0; q < kMaxPieces; q++)
                vq[q] = (pb + q < pe) ? __ldcg(&sm.piece_top[((size_t)(pb + q) * LFR + j) * 32u + lane]) : 0u;
            uint32_t rtop = 0;
#pragma unroll
            for (uint32_t q = 0; q < kMaxPieces; q++) {
                const uint32_t thr = __shfl_sync(0xFFFFFFFFu, rtop, (int)Kp - 1);
                unsigned rest = __ballot_sync(0xFFFFFFFFu, lane < Kp && vq[q] > thr);
                while (rest) {
                    const int src = __ffs(rest) - 1;
                    rest &= rest - 1;
                    rtop = lane_list_insert(rtop, __shfl_sync(0xFFFFFFFFu, vq[q], src), Kp);
                }
            }
            const uint32_t seed = __shfl_sync(0xFFFFFFFFu, rtop, (int)Kp - 1);
            if (lane == 0) sm.theta_seed[p * LFR + j] = seed;
        }
    }
    if (lane == 0) sm.ticket[p] = 0;
}

// Dynamic shared memory: bscsr_stream_smem(XREP, THREADS) bytes = XREP copies of the query + the ptab columns.
__host__ __device__ constexpr size_t bscsr_stream_smem(int xrep, int threads) {
    return (size_t)1024 * xrep * 4 + (size_t)16 * threads * 4;
}

template <int W, int LFR, int XREP, int THREADS, bool PREFETCH, bool BSX>
__global__ void __launch_bounds__(THREADS, (XREP <= 2 && THREADS <= 256) ? 4 : 1)
bscsr_stream_kernel(const uint8_t *__restrict__ packets, BscsrChunks m, const uint32_t *__restrict__ xq, uint32_t Kp,
                    BscsrLogs logs, const uint32_t *__restrict__ theta_seed, const uint32_t *__restrict__ sample_end,
                    uint32_t *chunk_counter) {
    extern __shared__ __align__(16) uint8_t bs_smem[];
    uint32_t *xs = reinterpret_cast<uint32_t *>(bs_smem);   // query, pre-shifted (bscsr_api.cu), XREP copies interleaved
    uint32_t *ptab = xs + 1024 * XREP;                      // [prefix length 0..15][thread]: running sums of the products
    pdl_trigger();   // the replay kernel's CTAs may be set up while this grid drains
    // the query words were complete before the sample kernel started: the 128 KB of copies are staged while it still runs
    for (uint32_t i = threadIdx.x; i < 1024u * XREP; i += THREADS) xs[i] = xq[i / XREP];
    ptab[threadIdx.x] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) chunk_counter[1] = 0;   // log-entry statistics of this run
    __syncthreads();
    pdl_wait();      // the sample's seeds (theta_seed) are not
    const unsigned lane = lane_id();
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(chunk_counter, 1u);
        c = __shfl_sync(0xFFFFFFFFu, c, 0);
        if (c >= m.n) break;
        const uint32_t first = m.first[c], count = m.count[c], local0 = m.local0[c];
        // chunks behind the sampled prefix of their partition start from the sample's K-th largest value
        const bool seeded = local0 >= sample_end[m.part[c]];
        uint32_t theta[LFR], top[LFR];
        BsLogSink<LFR> sink;
        sink.val = logs.val + (size_t)c * LFR * m.cap;
        sink.row = logs.row + (size_t)c * LFR * m.cap;
        sink.cap = m.cap;
#pragma unroll
        for (int j = 0; j < LFR; j++) {
            theta[j] = seeded ? theta_seed[m.part[c] * LFR + j] : 0u;
            top[j] = theta[j];   // as if K candidates of that value had been seen: max(seed, own K-th largest)
            sink.lcnt[j] = 0;
        }
        bs_run<W, LFR, XREP, THREADS, PREFETCH, BSX>(packets, first - m.lookback[c], first, first + count, local0,
                                                     m.row_in[c], Kp, reinterpret_cast<const uint8_t *>(xs), ptab, theta, top,
                                                     sink, logs.p0 + (size_t)c * LFR);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < LFR; j++) logs.cnt[(size_t)c * LFR + j] = sink.lcnt[j];
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

template <int W>
__global__ void __launch_bounds__(kReplayThreads)
bscsr_replay_kernel(BscsrLogs logs, const uint32_t *__restrict__ part_chunk_begin, uint32_t LFR, uint32_t Kp,
                    uint32_t chunk_cap, uint32_t *res_idx_words, uint32_t *res_val_words, uint32_t *chunk_counter_reset) {
    const uint32_t p = blockIdx.x / LFR, j = blockIdx.x % LFR;
    const uint32_t cb = part_chunk_begin[p], ce = part_chunk_begin[p + 1];
    const uint32_t tid = threadIdx.x;
    const unsigned lane = lane_id();
    extern __shared__ __align__(16) uint8_t replay_smem[];   // kReplayDynSmem bytes
    uint32_t *s_sv = reinterpret_cast<uint32_t *>(replay_smem), *s_sr = s_sv + kReplaySurvivors;
    __shared__ uint32_t s_wsum[kReplayThreads / 32], s_off[kReplayThreads * kReplayCpt + 1];
    __shared__ uint32_t s_n;
    if (tid == 0) s_n = 0;
    pdl_wait();      // the logs belong to the stream kernel right before
    if (blockIdx.x == 0 && tid == 0 && chunk_counter_reset) *chunk_counter_reset = 0;   // [1]: log entries, statistics
    const bool first_from_packet0 = logs.p0[(size_t)cb * LFR + j] != 0;
    __syncthreads();

    // Literal replace-min (hpp:366-389) over the buffered entries, by warp 0.  The K slots live one per lane
    // (lv/li of lanes 0..Kp-1), so the argmin is one warp reduction instead of a serial scan.  Lanes screen 32
    // entries at a time against the CURRENT worst value (it only grows, so an entry below it can never be
    // accepted); the rare accepted entries are applied one by one, in stream order.
    // (A per-lane register copy of the whole list with straight-line argmin measured slower: 41 us vs 28 us.)
    uint32_t lv = 0, li = 0;   // warp 0 only: value / row index of slot `lane`
    uint32_t wi = 0, wv = 0, started = 0;
    auto warp_argmin = [&](uint32_t &idx, uint32_t &val) {
        // MIN(res,a,b) = res[a] < res[b] ? a : b  -> the HIGHEST slot among equal minima (hpp:28);
        // K == 4 reproduces `MIN(res, 2, 2)` (hpp:45): slot 3 is never the minimum
        const bool in = lane < Kp && !(Kp == 4 && lane == 3);
        if (W <= 27) {
            // one reduction: key = value * 32 + (31 - slot); the smallest key is the smallest value in the highest slot
            const uint32_t key = in ? ((lv << 5) | (31u - lane)) : 0xFFFFFFFFu;
            const uint32_t mk = __reduce_min_sync(0xFFFFFFFFu, key);
            idx = 31u - (mk & 31u);
            val = mk >> 5;
        } else {
            const uint32_t mn = __reduce_min_sync(0xFFFFFFFFu, in ? lv : 0xFFFFFFFFu);
            const unsigned who = __ballot_sync(0xFFFFFFFFu, in && lv == mn);
            idx = 31u - (uint32_t)__clz((int)who);
            val = mn;
        }
    };
    auto replay = [&]() {
        if (tid < 32) {
            const uint32_t n = s_n;
            for (uint32_t b = 0; b < n; b += 32) {
                const uint32_t i = b + lane;
                const uint32_t v = (i < n) ? s_sv[i] : 0u, r = (i < n) ? s_sr[i] : 0u;
                if (!started) {
                    // the argmin is recomputed after EVERY packet (hpp:376-388): unless the first candidate comes
                    // from packet 0 of the partition, the all-zero list has already moved the worst slot
                    if (!first_from_packet0) warp_argmin(wi, wv);
                    started = 1;
                }
                unsigned rest = __ballot_sync(0xFFFFFFFFu, i < n && v >= wv);
                while (rest) {
                    const int src = __ffs(rest) - 1;
                    rest &= rest - 1;
                    const uint32_t cv = __shfl_sync(0xFFFFFFFFu, v, src), cr = __shfl_sync(0xFFFFFFFFu, r, src);
                    if (cv >= wv) {   // warp-uniform
                        if (lane == wi) { li = cr; lv = cv; }
                        warp_argmin(wi, wv);
                    }
                }
            }
            __syncwarp();   // every lane has read s_n (racecheck: the loop's votes order execution, not this read against the write)
            if (lane == 0) s_n = 0;
        }
        __syncthreads();
    };

    // chunks of the partition, kReplayThreads * kReplayCpt at a time (one tile covers any cfg3 partition): thread t owns
    // kReplayCpt consecutive chunks; the tile's log entries are copied to the buffer in stream order
    const uint32_t tile = blockDim.x * kReplayCpt;
    for (uint32_t t0 = cb; t0 < ce; t0 += tile) {
        uint32_t cnt[kReplayCpt], mine = 0;
#pragma unroll
        for (uint32_t u = 0; u < kReplayCpt; u++) {
            const uint32_t c = t0 + tid * kReplayCpt + u;
            cnt[u] = (c < ce) ? logs.cnt[(size_t)c * LFR + j] : 0u;
            mine += cnt[u];
        }
        // block-wide exclusive scan of the per-thread totals
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((int)lane >= d) inc += up;
        }
        if (lane == 31) s_wsum[tid / 32] = inc;
        __syncthreads();
        uint32_t before = 0, all = 0;
        for (uint32_t wv = 0; wv < blockDim.x / 32; wv++) { if (wv < tid / 32) before += s_wsum[wv]; all += s_wsum[wv]; }
        if (tid == 0 && chunk_counter_reset) atomicAdd(chunk_counter_reset + 1, all);
        if (s_n + all > kReplaySurvivors) {   // uniform: flush what is buffered first
            __syncthreads();
            replay();
        }
        if (all > kReplaySurvivors) {
            // logs longer than the buffer (adversarial inputs): replay them straight from memory, in order
            const uint32_t tend = (t0 + tile < ce) ? t0 + tile : ce;
            for (uint32_t cc = t0; cc < tend; cc++) {
                const uint32_t n_c = logs.cnt[(size_t)cc * LFR + j];
                const size_t lbc = ((size_t)cc * LFR + j) * chunk_cap;
                for (uint32_t e0 = 0; e0 < n_c; e0 += kReplaySurvivors) {
                    const uint32_t ne = (n_c - e0 < kReplaySurvivors) ? (n_c - e0) : kReplaySurvivors;
                    for (uint32_t e = tid; e < ne; e += blockDim.x) { s_sv[e] = logs.val[lbc + e0 + e]; s_sr[e] = logs.row[lbc + e0 + e]; }
                    if (tid == 0) s_n = ne;
                    __syncthreads();
                    replay();
                }
            }
            continue;
        }
        // order-preserving gather: entry f of the tile's concatenated logs belongs to the chunk whose
        // exclusive offset is the last one <= f
        {
            uint32_t o = before + inc - mine;
#pragma unroll
            for (uint32_t u = 0; u < kReplayCpt; u++) { s_off[tid * kReplayCpt + u] = o; o += cnt[u]; }
        }
        if (tid == 0) s_off[tile] = all;
        __syncthreads();
        const uint32_t base_n = s_n;
        for (uint32_t f = tid; f < all; f += blockDim.x) {
            uint32_t lo = 0, hi = tile;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_off[mid] <= f) lo = mid; else hi = mid; }
            const size_t e = ((size_t)(t0 + lo) * LFR + j) * chunk_cap + (f - s_off[lo]);
            s_sv[base_n + f] = logs.val[e];
            s_sr[base_n + f] = logs.row[e];
        }
        __syncthreads();
        if (tid == 0) s_n = base_n + all;
        __syncthreads();
    }
    replay();
    // write-back (.cpp:151-185): word t, position j = list j slot t; values widened to ufixed<32,1>
    if (tid < Kp) {
        const size_t o = ((size_t)p * Kp + tid) * 16u + j;
        res_idx_words[o] = li;
        res_val_words[o] = (W == 32) ? lv : (lv << (32 - W));
    }
}

}  // namespace tks
