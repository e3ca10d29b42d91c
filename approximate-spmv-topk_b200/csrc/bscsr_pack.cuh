// bscsr_pack.cuh -- GPU-side BS-CSR packer.
//
// The reference builds its packets with one host thread per run (src/fpga/src/host_spmv_bscsr.cpp:112-121
// row partitioning, :133-187 packet_coo, :189-248 packet_coo_partition, fpga_utils.hpp:307-365 bit layout);
// at 10^7 rows that and the device tables derived from the packets (bscsr_api.cu: row counter and carry
// look-back at every chunk start, the BSX re-encoding) cost seconds of setup.  Here all of it runs on the
// device from the row-sorted COO the reference constructor takes (host_spmv_bscsr.cpp:104):
//
//   bscsr_pack_check_kernel     rows sorted, row < num_rows, col < cols
//   bscsr_partition_kernel      first non-zero of every row range [p * rpp, (p+1) * rpp)     (host:136-141)
//   bscsr_pack_kernel<W>        one thread per packet: segment ends, xf, quantised values; writes the
//                               reference's 512-bit word or the BSX words, plus the packet's row-counter
//                               advance and its "passes the carry through" flag
//   (exclusive scan of the advances: csr_build.cuh's scan kernels)
//   bscsr_chunk_walk_kernel     one thread per partition: chunk and sample-piece boundaries (each depends on
//                               the look-back of the one before, a short sequential walk per partition)
//   bscsr_patch_rel_kernel      BSX: row counter of every packet relative to its chunk -> word 15
//
// The resident state is identical, byte for byte, to tks_pack_bscsr + tks_upload_bscsr (tests/test_gpu_pack.py).
#pragma once

#include "bscsr_topk.cuh"

namespace tks {

constexpr uint32_t kPackErrUnsorted = 1u, kPackErrRowRange = 2u, kPackErrColRange = 4u, kPackErrRel = 8u;

__global__ void bscsr_pack_check_kernel(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col, uint64_t nnz,
                                        uint32_t num_rows, uint32_t cols, uint32_t *err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t bad = 0;
    for (; i < nnz; i += stride) {
        const uint32_t r = row[i];
        if (i + 1 < nnz && row[i + 1] < r) bad |= kPackErrUnsorted;
        if (r >= num_rows) bad |= kPackErrRowRange;
        if (col[i] >= cols) bad |= kPackErrColRange;
    }
    if (bad) atomicOr(err, bad);
}

// nnz_start[p] = first i with row[i] >= p * rpp (p = 0..P), by binary search.
__global__ void bscsr_partition_kernel(const uint32_t *__restrict__ row, uint64_t nnz, uint32_t rpp, uint32_t P,
                                       uint64_t *nnz_start) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > P) return;
    const uint64_t lim = (uint64_t)rpp * p;
    uint64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((uint64_t)row[mid] < lim) lo = mid + 1; else hi = mid;
    }
    nnz_start[p] = (p == P) ? nnz : lo;
}

struct PackParts {
    const uint64_t *nnz_start;    // [P+1]
    const uint64_t *pkt_start;    // [P+1] global index of every partition's first packet
    uint32_t P;
};

// ap_ufixed<32,1>::to_float() (nearest even) then (ap_ufixed<W,1,AP_TRN_ZERO>) float: fpga_utils.hpp:336-338
__device__ __forceinline__ uint32_t fixedW_from_fixed32_dev(uint32_t raw32, int W) {
    const float f = __uint2float_rn(raw32) * 4.656612873077393e-10f;              // * 2^-31, exact
    const unsigned long long t = __float2ull_rz(f * (float)(1ull << (W - 1)));    // power of two: exact, then truncate
    return (uint32_t)(t & ((W == 32) ? 0xFFFFFFFFull : ((1ull << W) - 1ull)));
}

__device__ __forceinline__ void put_bits(uint64_t (&w)[8], unsigned lo, unsigned width, uint64_t v) {
    v &= (1ull << width) - 1ull;   // width <= 32
    const unsigned q = lo >> 6, s = lo & 63;
    w[q] |= v << s;
    if (s + width > 64) w[q + 1] |= v >> (64 - s);
}

// One thread per packet.  BSX: the re-encoded device words of bscsr_topk.cuh (row-counter field patched later);
// otherwise the reference's packet, bit for bit.
template <int W, bool BSX>
__global__ void bscsr_pack_kernel(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col,
                                  const uint32_t *__restrict__ val32, PackParts parts, uint64_t total_packets,
                                  int LFR, int drift_free, uint8_t *__restrict__ packets, uint32_t *__restrict__ advance,
                                  uint8_t *__restrict__ keep) {
    constexpr int B = 511 / (W + 14);
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_packets) return;
    // partition of the packet
    uint32_t lo = 0, hi = parts.P;   // largest p with pkt_start[p] <= g
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (parts.pkt_start[mid] <= g) lo = mid; else hi = mid;
    }
    const uint32_t p = lo;
    const uint64_t i = g - parts.pkt_start[p];
    const uint64_t ns = parts.nnz_start[p], nnz_p = parts.nnz_start[p + 1] - ns;
    const uint64_t base = ns + i * (uint64_t)B;
    const int cnt = (int)((nnz_p - i * (uint64_t)B < (uint64_t)B) ? (nnz_p - i * (uint64_t)B) : (uint64_t)B);
    // row of the last entry before the packet: the previous entry of the stream; for the very first packet 0
    // (host:153-157: partition p starts from the last row of partition p-1, which is the previous entry)
    const uint32_t prev_row = (base == 0) ? 0u : row[base - 1];
    uint32_t r[B], ends[B];
#pragma unroll
    for (int j = 0; j < B; j++) r[j] = (j < cnt) ? row[base + j] : 0u;
    const uint32_t xf = r[0] != prev_row;
    int nseg = 0;
#pragma unroll
    for (int j = 0; j < B; j++) {
        if (j < cnt) {
            const bool e = (j + 1 == cnt) || (r[j + 1] != r[j]);
            if (e) ends[nseg++] = (uint32_t)(j + 1);
        }
    }
    const uint32_t last_end = ends[nseg - 1];
    const uint32_t nw = (i != 0) ? xf : 0u;
    const uint32_t n = (uint32_t)(nseg < LFR ? nseg : LFR);     // non-empty segments among the first LFR (hpp:131-142)
    const bool over = drift_free && nseg > LFR;
    uint32_t zero_from = 0, zero_to = 0;
    uint64_t xw = 0;   // 4-bit cumulative ends
#pragma unroll
    for (int s = 0; s < B; s++) {
        uint32_t e = (s < nseg) ? ends[s] : last_end;
        if (over) e = (s < LFR - 1) ? ends[s] : last_end;
        xw |= (uint64_t)e << (4 * s);
    }
    if (over) {
        // drift-free re-encoding (bscsr_api.cu): 16 nibbles, the tail repeats the last end
        for (int s = B; s < 16; s++) xw |= (uint64_t)last_end << (4 * s);
        zero_from = ends[LFR - 2];
        zero_to = ends[nseg - 2];
    }
    advance[g] = (over ? (uint32_t)nseg : n) + nw - 1u;
    keep[g] = (n == 1 && nw == 0) || (n == 0 && nw != 0);
    if (BSX) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j++) w[j] = 0u;
#pragma unroll
        for (int j = 0; j < B; j++) {
            if (j < cnt) {
                const uint32_t v = ((uint32_t)j >= zero_from && (uint32_t)j < zero_to) ? 0u : fixedW_from_fixed32_dev(val32[base + j], W);
                w[j] = (v << (32 - W)) | (col[base + j] & 0x3FFu);
            }
        }
        w[15] = (uint32_t)(xw & 0xFFFFu) | ((nw | (n << 1)) << 16);
        uint4 *out = reinterpret_cast<uint4 *>(packets + g * 64u);
        out[0] = make_uint4(w[0], w[1], w[2], w[3]);
        out[1] = make_uint4(w[4], w[5], w[6], w[7]);
        out[2] = make_uint4(w[8], w[9], w[10], w[11]);
        out[3] = make_uint4(w[12], w[13], w[14], w[15]);
    } else {
        uint64_t w[8];
#pragma unroll
        for (int q = 0; q < 8; q++) w[q] = 0ull;
        put_bits(w, 511, 1, xf);
        w[0] |= xw & ((B * 4 >= 64) ? ~0ull : ((1ull << (B * 4)) - 1ull));
        for (int j = 0; j < cnt; j++) {
            put_bits(w, 4u * B + 10u * (unsigned)j, 10, col[base + j]);
            put_bits(w, 14u * B + (unsigned)W * (unsigned)j, (unsigned)W, fixedW_from_fixed32_dev(val32[base + j], W));
        }
        uint4 *out = reinterpret_cast<uint4 *>(packets + g * 64u);
        out[0] = make_uint4((uint32_t)w[0], (uint32_t)(w[0] >> 32), (uint32_t)w[1], (uint32_t)(w[1] >> 32));
        out[1] = make_uint4((uint32_t)w[2], (uint32_t)(w[2] >> 32), (uint32_t)w[3], (uint32_t)(w[3] >> 32));
        out[2] = make_uint4((uint32_t)w[4], (uint32_t)(w[4] >> 32), (uint32_t)w[5], (uint32_t)(w[5] >> 32));
        out[3] = make_uint4((uint32_t)w[6], (uint32_t)(w[6] >> 32), (uint32_t)w[7], (uint32_t)(w[7] >> 32));
    }
}

// Chunk / sample-piece tables of one partition (same walk as the host path in bscsr_upload).
struct WalkOut {
    uint32_t *first, *count, *local0, *row_in, *look, *part;   // chunk tables (nullptr in the counting pass)
    uint32_t *s_first, *s_count, *s_local0, *s_look, *s_part;  // sample pieces
    uint32_t *sample_end;                                      // [P]
    uint32_t *n_chunks, *n_pieces;                             // [P] counts (counting pass)
    const uint32_t *chunk_begin, *piece_begin;                 // [P] output offsets (fill pass)
};

__device__ __forceinline__ uint32_t pack_lookback(const uint8_t *keep, uint64_t i) {
    if (i == 0) return 0u;
    uint32_t L = 1;
    while (i - L > 0 && keep[i - L]) L++;
    return L;
}

// rowsum: exclusive scan of `advance` over ALL packets (u64); the kernel's row counter before packet g of
// partition p is rowsum[g] - rowsum[pkt_start[p]].
__global__ void bscsr_chunk_walk_kernel(PackParts parts, const uint8_t *__restrict__ keep, const uint64_t *__restrict__ rowsum,
                                        uint64_t total_packets, uint32_t chunk_cap, uint32_t tail_div, WalkOut o) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= parts.P) return;
    const uint64_t g0 = parts.pkt_start[p], np = parts.pkt_start[p + 1] - g0;
    const uint8_t *kp = keep + g0;
    const uint64_t tail_begin = total_packets - total_packets / 10u;
    const bool fill = o.first != nullptr;
    uint32_t nc = 0, cb = fill ? o.chunk_begin[p] : 0u;
    for (uint64_t i = 0; i < np;) {
        const uint32_t L = pack_lookback(kp, i);
        const uint32_t cap_here = (g0 + i >= tail_begin) ? chunk_cap / tail_div : chunk_cap;
        const uint64_t room = cap_here - (L % 32u);
        const uint64_t cnt = room < np - i ? room : np - i;
        if (fill) {
            o.first[cb + nc] = (uint32_t)(g0 + i);
            o.count[cb + nc] = (uint32_t)cnt;
            o.local0[cb + nc] = (uint32_t)i;
            o.row_in[cb + nc] = (uint32_t)(rowsum[g0 + i] - rowsum[g0]);
            o.look[cb + nc] = L;
            o.part[cb + nc] = p;
        }
        nc++;
        i += cnt;
    }
    uint32_t npc = 0, pb = fill ? o.piece_begin[p] : 0u;
    uint64_t next_piece = 0;
    const uint64_t lim = np < kBsSamplePackets ? np : (uint64_t)kBsSamplePackets;
    while (next_piece < lim && npc < kBsMaxPieces) {
        const uint64_t i = next_piece;
        const uint32_t L = pack_lookback(kp, i);
        const uint64_t room = kBsSamplePiece - (L % 32u);
        const uint64_t cnt = room < lim - i ? room : lim - i;
        if (fill) {
            o.s_first[pb + npc] = (uint32_t)(g0 + i);
            o.s_count[pb + npc] = (uint32_t)cnt;
            o.s_local0[pb + npc] = (uint32_t)i;
            o.s_look[pb + npc] = L;
            o.s_part[pb + npc] = p;
        }
        npc++;
        next_piece = i + cnt;
    }
    if (fill) o.sample_end[p] = (uint32_t)next_piece;
    else { o.n_chunks[p] = nc; o.n_pieces[p] = npc; }
}

// BSX word 15, bits 20..31: row counter before the packet relative to its chunk, + nw (bscsr_api.cu / bscsr_topk.cuh).
// One warp per chunk.
__global__ void bscsr_patch_rel_kernel(uint8_t *__restrict__ packets, const uint32_t *__restrict__ chunk_first,
                                       const uint32_t *__restrict__ chunk_count, const uint32_t *__restrict__ chunk_part,
                                       const uint32_t *__restrict__ chunk_row_in, uint32_t n_chunks,
                                       const uint64_t *__restrict__ rowsum, const uint64_t *__restrict__ pkt_start,
                                       uint32_t *err) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    if (c >= n_chunks) return;
    const uint32_t first = chunk_first[c], cnt = chunk_count[c];
    const uint64_t part_sum = rowsum[pkt_start[chunk_part[c]]];
    const uint32_t row_in = chunk_row_in[c];
    for (uint32_t t = lane_id(); t < cnt; t += kWarp) {
        const uint64_t g = (uint64_t)first + t;
        uint32_t *w15 = reinterpret_cast<uint32_t *>(packets + g * 64u) + 15;
        const uint32_t w = *w15;
        const uint32_t nw = (w >> 16) & 1u;
        const uint32_t rel = (uint32_t)(rowsum[g] - part_sum) - row_in + nw;
        if (rel > 0xFFFu) atomicOr(err, kPackErrRel);
        *w15 = w | (rel << 20);
    }
}

}  // namespace tks
