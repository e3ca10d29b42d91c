// host_api.cpp -- the CPU-only entry points of include/topkspmv.h: the reference's host surface
// around the accelerator (MTX loader, COO->CSR, value quantisation, BS-CSR packet builder), exported
// through the same C ABI so that non-C++ callers (ctypes tests, bench.py) use the very code the
// host executable uses.  No CUDA here and no top-k computation: nothing in this file is a fallback.
#include <algorithm>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/topkspmv.h"
#include "../host/bscsr_packer.hpp"
#include "../host/fixed_point.hpp"
#include "../host/matrix_cache.hpp"
#include "../host/mtx_reader.hpp"

static thread_local std::string g_host_error;

extern "C" {

const char *tks_host_last_error(void) { return g_host_error.c_str(); }

int tks_bscsr_packet_size(int fixed_width) {
    if (fixed_width < 1 || fixed_width > 64) return TKS_EINVAL;
    return tkshost::bscsr_packet_size(fixed_width);
}

uint32_t tks_fixed32_from_double(double v) { return ufixed32::from_double(v); }

uint32_t tks_fixedW_from_fixed32(uint32_t raw32, int fixed_width) {
    return tkshost::fixedW_from_fixed32(raw32, fixed_width);
}

// read_result (host_spmv_bscsr.cpp:399-448): all P x Kp x B slots, idx += first_row[p], keep val > 0, the first
// insertion of an index wins, then sort_tuples (evaluation_utils.hpp:40-62).  At most P x Kp x LFR candidates:
// a sort replaces the reference's unordered_map; same outcome.
int tks_merge_partition_words(uint32_t partitions, uint32_t local_k, uint32_t packet_size, const uint32_t *idx_words,
                              const uint32_t *val_words, const uint32_t *first_row, int tie_break, uint32_t k,
                              uint32_t *idx_out, uint32_t *val_out, uint32_t *count) {
    if (!idx_words || !val_words || !first_row || !count) { g_host_error = "null argument"; return TKS_EINVAL; }
    if (packet_size == 0 || packet_size > 16) { g_host_error = "packet_size outside 1..16"; return TKS_EINVAL; }
    // candidates in insertion order; an open-addressing table keeps the first insertion of every index (the
    // reference's unordered_map::insert, host:415-430); one 64-bit key per survivor orders them like sort_tuples:
    // value descending, then index descending (TKS_TIE_HIGHER_INDEX) or ascending
    static thread_local std::vector<uint64_t> keys, cand;
    static thread_local std::vector<uint32_t> table;
    const bool higher = tie_break == TKS_TIE_HIGHER_INDEX;
    cand.clear();
    for (uint32_t p = 0; p < partitions; p++) {
        const uint32_t base_row = first_row[p];
        const uint32_t *iw = idx_words + (size_t)p * local_k * 16, *vw = val_words + (size_t)p * local_k * 16;
        for (uint32_t t = 0; t < local_k; t++, iw += 16, vw += 16)
            for (uint32_t q = 0; q < packet_size; q++)
                if (vw[q] != 0) cand.push_back(((uint64_t)vw[q] << 32) | (uint32_t)(iw[q] + base_row));
    }
    size_t tsize = 64;
    while (tsize < 2 * cand.size()) tsize <<= 1;
    table.assign(tsize, 0u);   // idx + 1, 0 = empty (idx + 1 == 0 only for idx = 0xFFFFFFFF, which no 32-bit row count produces)
    keys.clear();
    for (const uint64_t c : cand) {
        const uint32_t idx = (uint32_t)c;
        size_t hpos = ((size_t)idx * 2654435761u >> 7) & (tsize - 1);
        bool seen = false;
        while (table[hpos] != 0) {
            if (table[hpos] == idx + 1u) { seen = true; break; }
            hpos = (hpos + 1) & (tsize - 1);
        }
        if (seen) continue;
        table[hpos] = idx + 1u;
        keys.push_back((c & 0xFFFFFFFF00000000ull) | (higher ? idx : ~idx));
    }
    *count = (uint32_t)keys.size();   // all distinct candidates; the caller's buffers receive the first min(k, count)
    if (idx_out && val_out) {
        const size_t want = k < keys.size() ? k : keys.size();
        auto desc = [](uint64_t l, uint64_t r) { return l > r; };
        if (want < keys.size()) {   // the k largest first (linear), then only those are ordered
            std::nth_element(keys.begin(), keys.begin() + (ptrdiff_t)want, keys.end(), desc);
            std::sort(keys.begin(), keys.begin() + (ptrdiff_t)want, desc);
        } else {
            std::sort(keys.begin(), keys.end(), desc);
        }
        for (uint32_t i = 0; i < k; i++) {
            const bool in = i < want;
            const uint32_t lo = in ? (uint32_t)keys[i] : 0u;
            idx_out[i] = in ? (higher ? lo : ~lo) : 0u;
            val_out[i] = in ? (uint32_t)(keys[i] >> 32) : 0u;
        }
    }
    return TKS_OK;
}

int tks_pack_bscsr(const uint32_t *row, const uint32_t *col, const uint32_t *val32, uint64_t nnz, uint32_t num_rows,
                   int partitions, int fixed_width, uint64_t *packets_per_part, uint32_t *first_row,
                   uint64_t *nnz_per_part, void *packets) {
    if (!row || !col || !val32 || !packets_per_part || !first_row || !nnz_per_part) { g_host_error = "null argument"; return TKS_EINVAL; }
    if (fixed_width < 17 || fixed_width > 32 || partitions < 1) { g_host_error = "fixed_width outside 17..32 or partitions < 1"; return TKS_EINVAL; }
    tkshost::BscsrPartitioning part;
    std::string err;
    int rc = tkshost::bscsr_partition(row, nnz, num_rows, partitions, fixed_width, part, &err);
    if (rc != 0) { g_host_error = err; return TKS_EINVAL; }
    for (int p = 0; p < partitions; p++) {
        packets_per_part[p] = part.num_packets[p];
        first_row[p] = part.first_row[p];
        nnz_per_part[p] = part.nnz_start[p + 1] - part.nnz_start[p];
    }
    if (!packets) return TKS_OK;
    auto *out = static_cast<tkshost::Packet512 *>(packets);
    uint64_t off = 0;
    for (int p = 0; p < partitions; p++) {
        const uint64_t s = part.nnz_start[p];
        tkshost::bscsr_pack_partition(row + s, col + s, val32 + s, part.nnz_start[p + 1] - s,
                                      p == 0 ? 0u : part.last_row[p - 1], fixed_width, out + off);
        off += part.num_packets[p];
    }
    return TKS_OK;
}

int tks_read_mtx(const char *path, int zero_indexed, int sort_tuples, int ignore_values, uint32_t *rows, uint32_t *cols,
                 uint64_t *nnz, uint64_t nnz_capacity, uint32_t *x, uint32_t *y, double *val) {
    if (!path || !rows || !cols || !nnz) { g_host_error = "null argument"; return TKS_EINVAL; }
    static thread_local std::string cached_path;
    static thread_local std::vector<uint32_t> cx, cy;
    static thread_local std::vector<double> cv;
    static thread_local uint32_t crows, ccols;
    static thread_local int cflags = -1;
    const int flags = (zero_indexed ? 1 : 0) | (sort_tuples ? 2 : 0) | (ignore_values ? 4 : 0);
    if (cached_path != path || cflags != flags) {
        uint32_t r = 0, c = 0, n = 0;
        std::string err;
        int rc = tkshost::readMtx<uint32_t, double>(path, &cx, &cy, &cv, &r, &c, &n, 0, !ignore_values, false,
                                                    zero_indexed != 0, sort_tuples != 0, &err);
        if (rc != 0) { g_host_error = err; cached_path.clear(); return TKS_EIO; }
        cached_path = path; cflags = flags; crows = r; ccols = c;
    }
    *rows = crows; *cols = ccols; *nnz = cx.size();
    if (nnz_capacity == 0) return TKS_OK;
    if (nnz_capacity < cx.size() || !x || !y || !val) { g_host_error = "output buffers too small"; return TKS_EINVAL; }
    std::memcpy(x, cx.data(), cx.size() * 4);
    std::memcpy(y, cy.data(), cy.size() * 4);
    std::memcpy(val, cv.data(), cv.size() * 8);
    cached_path.clear(); cx.clear(); cx.shrink_to_fit(); cy.clear(); cy.shrink_to_fit(); cv.clear(); cv.shrink_to_fit();
    return TKS_OK;
}

int tks_coo2csr(const uint32_t *x, const uint32_t *y, const float *val, uint64_t nnz, uint32_t rows, uint32_t cols,
                uint32_t *ptr, uint32_t *idx, float *out_val) {
    if (!x || !y || !val || !ptr || !idx || !out_val) { g_host_error = "null argument"; return TKS_EINVAL; }
    std::vector<uint32_t> xs(x, x + nnz), ys(y, y + nnz);
    std::vector<float> vs(val, val + nnz);
    if (tkshost::coo2csr<uint32_t, float>(ptr, idx, out_val, xs, ys, vs, rows, cols) != 0) {
        g_host_error = "Error: Index out of bounds!";
        return TKS_EINVAL;
    }
    return TKS_OK;
}

uint64_t tks_cache_source_tag(const char *source_path, int zero_indexed, int ignore_values) {
    return source_path ? tkshost::cache_source_tag(source_path, zero_indexed, ignore_values) : 0;
}

int tks_cache_write_csr(const char *path, uint64_t rows, uint32_t cols, uint64_t nnz, const uint64_t *ptr64,
                        const uint32_t *idx, const float *val) {
    return tks_cache_write_csr_tagged(path, rows, cols, nnz, ptr64, idx, val, 0);
}

int tks_cache_write_csr_tagged(const char *path, uint64_t rows, uint32_t cols, uint64_t nnz, const uint64_t *ptr64,
                               const uint32_t *idx, const float *val, uint64_t source_tag) {
    if (!path || !ptr64 || (nnz && (!idx || !val))) { g_host_error = "null argument"; return TKS_EINVAL; }
    if (ptr64[rows] != nnz) { g_host_error = "ptr[rows] != nnz"; return TKS_EINVAL; }
    tkshost::CacheHeader h{};
    h.kind = tkshost::kCacheCsr; h.cols = cols; h.rows = rows; h.nnz = nnz; h.aux0 = source_tag;
    std::string err;
    if (tkshost::cache_write(path, h, {{ptr64, (rows + 1) * 8}, {idx, nnz * 4}, {val, nnz * 4}}, &err) != 0) {
        g_host_error = err; return TKS_EIO;
    }
    return TKS_OK;
}

int tks_cache_read_csr(const char *path, uint64_t *rows, uint32_t *cols, uint64_t *nnz, uint64_t *ptr64, uint32_t *idx,
                       float *val) {
    return tks_cache_read_csr_tagged(path, rows, cols, nnz, ptr64, idx, val, nullptr);
}

int tks_cache_read_csr_tagged(const char *path, uint64_t *rows, uint32_t *cols, uint64_t *nnz, uint64_t *ptr64,
                              uint32_t *idx, float *val, uint64_t *source_tag) {
    if (!path || !rows || !cols || !nnz) { g_host_error = "null argument"; return TKS_EINVAL; }
    tkshost::CacheHeader h{};
    FILE *f = nullptr;
    std::string err;
    if (tkshost::cache_open(path, tkshost::kCacheCsr, &h, &f, &err) != 0) { g_host_error = err; return TKS_EIO; }
    if (h.payload_bytes != (h.rows + 1) * 8 + h.nnz * 8) { std::fclose(f); g_host_error = std::string(path) + ": inconsistent header"; return TKS_EIO; }
    *rows = h.rows; *cols = h.cols; *nnz = h.nnz;
    if (source_tag) *source_tag = h.aux0;
    if (!ptr64 && !idx && !val) { std::fclose(f); return TKS_OK; }          // size query
    if (!ptr64 || (h.nnz && (!idx || !val))) { std::fclose(f); g_host_error = "null output array"; return TKS_EINVAL; }
    if (tkshost::cache_read_sections(f, h, {{ptr64, (h.rows + 1) * 8}, {idx, h.nnz * 4}, {val, h.nnz * 4}}, path, &err) != 0) {
        g_host_error = err; return TKS_EIO;
    }
    if (ptr64[h.rows] != h.nnz) { g_host_error = std::string(path) + ": ptr[rows] != nnz"; return TKS_EIO; }
    return TKS_OK;
}

int tks_cache_write_bscsr(const char *path, uint32_t rows, uint32_t cols, int fixed_width, uint32_t partitions,
                          const uint64_t *packets_per_part, const uint32_t *first_row, const uint64_t *nnz_per_part,
                          const void *packets) {
    if (!path || !packets_per_part || !first_row || !nnz_per_part || !packets) { g_host_error = "null argument"; return TKS_EINVAL; }
    uint64_t total = 0, nnz = 0;
    for (uint32_t p = 0; p < partitions; p++) { total += packets_per_part[p]; nnz += nnz_per_part[p]; }
    tkshost::CacheHeader h{};
    h.kind = tkshost::kCacheBscsr; h.cols = cols; h.rows = rows; h.nnz = nnz; h.aux0 = partitions; h.aux1 = (uint64_t)fixed_width;
    std::string err;
    if (tkshost::cache_write(path, h, {{packets_per_part, (size_t)partitions * 8}, {first_row, (size_t)partitions * 4},
                                       {nnz_per_part, (size_t)partitions * 8}, {packets, total * 64}}, &err) != 0) {
        g_host_error = err; return TKS_EIO;
    }
    return TKS_OK;
}

int tks_cache_read_bscsr(const char *path, uint32_t *rows, uint32_t *cols, int *fixed_width, uint32_t *partitions,
                         uint64_t *total_packets, uint64_t *packets_per_part, uint32_t *first_row,
                         uint64_t *nnz_per_part, void *packets) {
    if (!path || !rows || !cols || !fixed_width || !partitions || !total_packets) { g_host_error = "null argument"; return TKS_EINVAL; }
    tkshost::CacheHeader h{};
    FILE *f = nullptr;
    std::string err;
    if (tkshost::cache_open(path, tkshost::kCacheBscsr, &h, &f, &err) != 0) { g_host_error = err; return TKS_EIO; }
    const uint64_t P = h.aux0;
    if (P == 0 || P > 4096 || h.payload_bytes < P * 20 || (h.payload_bytes - P * 20) % 64 != 0) {
        std::fclose(f); g_host_error = std::string(path) + ": inconsistent header"; return TKS_EIO;
    }
    *rows = (uint32_t)h.rows; *cols = h.cols; *fixed_width = (int)h.aux1; *partitions = (uint32_t)P;
    *total_packets = (h.payload_bytes - P * 20) / 64;
    if (!packets_per_part && !first_row && !nnz_per_part && !packets) { std::fclose(f); return TKS_OK; }   // size query
    if (!packets_per_part || !first_row || !nnz_per_part || !packets) { std::fclose(f); g_host_error = "null output array"; return TKS_EINVAL; }
    if (tkshost::cache_read_sections(f, h, {{packets_per_part, P * 8}, {first_row, P * 4}, {nnz_per_part, P * 8},
                                            {packets, *total_packets * 64}}, path, &err) != 0) {
        g_host_error = err; return TKS_EIO;
    }
    uint64_t total = 0;
    for (uint64_t p = 0; p < P; p++) total += packets_per_part[p];
    if (total != *total_packets) { g_host_error = std::string(path) + ": packet counts do not add up"; return TKS_EIO; }
    return TKS_OK;
}

}  // extern "C"
