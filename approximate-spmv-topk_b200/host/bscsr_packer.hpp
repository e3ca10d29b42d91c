// bscsr_packer.hpp -- row partitioning and BS-CSR packet builder of the host surface.
//
// Produces, bit for bit, the 512-bit packets the reference host feeds its FPGA kernel:
//   partitioning        src/fpga/src/host_spmv_bscsr.cpp:112-121, 136-150
//   packet_coo_partition src/fpga/src/host_spmv_bscsr.cpp:189-248
//   bit layout          src/fpga/src/ip/fpga_utils.hpp:307-365 (write_block_x/y/val/xf)
//   query packing       src/fpga/src/ip/fpga_utils.hpp:346-355, host_spmv_bscsr.cpp:173-186
// Packet (W = FIXED_WIDTH, B = floor(511 / (W + 14)) entries):
//   bits [0, 4B)        B x 4-bit  cumulative end offsets of the row segments inside the packet
//   bits [4B, 14B)      B x 10-bit column indices
//   bits [14B, 14B+WB)  B x W-bit  raw ap_ufixed<W,1> values
//   bit  511            xf: the first entry belongs to a different row than the previous packet's last
// Unlike the reference (one bitstream per W) W, B and the partition count are runtime parameters.
// The structure differs from the reference's (segment ends are emitted while walking the entries
// once, words are assembled with 64-bit shifts); tests/test_packer.py checks equality with the
// oracle's literal transcription and, where built, with the reference's own code.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace tkshost {

inline int bscsr_packet_size(int W) { return (512 - 1) / (W + 14); }   // types.hpp:71-72

inline uint32_t fixedW_from_fixed32(uint32_t raw32, int W) {
    // ap_ufixed<32,1>::to_float() (nearest even) then (ap_ufixed<W,1,AP_TRN_ZERO>) float: fpga_utils.hpp:336-338
    float f = (float)((double)raw32 / 2147483648.0);
    double s = (double)f * (double)(1ull << (W - 1));
    uint64_t t = (uint64_t)s;   // s >= 0: truncation == floor
    uint64_t m = (W == 32) ? 0xFFFFFFFFull : ((1ull << W) - 1ull);
    return (uint32_t)(t & m);
}

struct Packet512 {
    uint64_t w[8];
    void clear() { std::memset(w, 0, sizeof w); }
    // OR `width` (<= 32) bits of v at bit offset lo; fields never overlap, so OR == assign
    void put(unsigned lo, unsigned width, uint64_t v) {
        v &= (width == 64) ? ~0ull : ((1ull << width) - 1ull);
        unsigned q = lo >> 6, s = lo & 63;
        w[q] |= v << s;
        if (s + width > 64) w[q + 1] |= v >> (64 - s);
    }
};
static_assert(sizeof(Packet512) == 64, "packet must be 64 bytes");

struct BscsrPartitioning {
    std::vector<uint64_t> nnz_start;     // P+1 offsets into the row-sorted COO
    std::vector<uint32_t> first_row;     // host:145
    std::vector<uint32_t> last_row;      // host:146
    std::vector<uint64_t> num_packets;   // host:148
};

// host_spmv_bscsr.cpp:136-150.  rows must be sorted; every partition must be non-empty.
inline int bscsr_partition(const uint32_t *row, uint64_t nnz, uint32_t num_rows, int P, int W,
                           BscsrPartitioning &out, std::string *err = nullptr) {
    const int B = bscsr_packet_size(W);
    const uint32_t rpp = (num_rows + (uint32_t)P - 1) / (uint32_t)P;
    out.nnz_start.assign((size_t)P + 1, 0);
    out.first_row.assign(P, 0); out.last_row.assign(P, 0); out.num_packets.assign(P, 0);
    if (rpp == 0) { if (err) *err = "matrix has no rows"; return -1; }
    uint64_t i = 0;
    for (int p = 0; p < P; p++) {
        out.nnz_start[p] = i;
        const uint64_t lim = (uint64_t)rpp * (uint64_t)(p + 1);
        while (i < nnz && (uint64_t)row[i] < lim) {
            if (i + 1 < nnz && row[i + 1] < row[i]) { if (err) *err = "COO rows are not sorted"; return -2; }
            i++;
        }
        const uint64_t n = i - out.nnz_start[p];
        if (n == 0) {
            if (err) *err = "partition " + std::to_string(p) + " has no non-zeros (the reference requires every row range to be populated)";
            return -1;
        }
        out.first_row[p] = row[out.nnz_start[p]];
        out.last_row[p] = row[i - 1];
        out.num_packets[p] = (n + (uint64_t)B - 1) / (uint64_t)B;
    }
    out.nnz_start[P] = i;
    if (i != nnz) { if (err) *err = "row index beyond num_rows"; return -2; }
    return 0;
}

// host_spmv_bscsr.cpp:189-248 for one partition.  val32: raw ap_ufixed<32,1>.
inline void bscsr_pack_partition(const uint32_t *row, const uint32_t *col, const uint32_t *val32, uint64_t nnz_p,
                                 uint32_t prev_last_row, int W, Packet512 *out) {
    const int B = bscsr_packet_size(W);
    const uint64_t npk = (nnz_p + (uint64_t)B - 1) / (uint64_t)B;
    const unsigned y_off = 4u * (unsigned)B, v_off = 14u * (unsigned)B;
    uint32_t prev_row = prev_last_row;   // row of the last entry before the packet
    for (uint64_t i = 0; i < npk; i++) {
        Packet512 pk;
        pk.clear();
        const uint64_t base = i * (uint64_t)B;
        const int cnt = (int)((nnz_p - base < (uint64_t)B) ? (nnz_p - base) : (uint64_t)B);   // entries in range
        pk.put(511, 1, row[base] != prev_row);
        // Walk the entries once: a segment ends after entry j when entry j+1 is out of range or in another row.
        // Ends are cumulative counts; unused slots repeat the last end (a run of length 0).
        int seg = 0;
        uint32_t last_end = 0;
        for (int j = 0; j < cnt; j++) {
            pk.put(y_off + 10u * (unsigned)j, 10, col[base + j]);
            pk.put(v_off + (unsigned)W * (unsigned)j, (unsigned)W, fixedW_from_fixed32(val32[base + j], W));
            const bool ends = (j + 1 == cnt) || (row[base + j + 1] != row[base + j]);
            if (ends) {
                last_end = (uint32_t)(j + 1);
                pk.put(4u * (unsigned)seg, 4, last_end);
                seg++;
            }
        }
        for (; seg < B; seg++) pk.put(4u * (unsigned)seg, 4, last_end);
        prev_row = row[base + cnt - 1];
        out[i] = pk;
    }
}

// write_block_vec + host:173-186: B raw 32-bit query words per 64-byte block.
inline void bscsr_pack_query(const uint32_t *vec32, uint32_t cols, int W, std::vector<Packet512> &out) {
    const int B = bscsr_packet_size(W);
    const uint32_t nblk = (cols + (uint32_t)B - 1) / (uint32_t)B;
    out.assign(nblk, Packet512{});
    for (uint32_t i = 0; i < nblk; i++) {
        out[i].clear();
        for (int j = 0; j < B; j++) {
            uint32_t c = i * (uint32_t)B + (uint32_t)j;
            out[i].put(32u * (unsigned)j, 32, c < cols ? vec32[c] : 0u);
        }
    }
}

}  // namespace tkshost
