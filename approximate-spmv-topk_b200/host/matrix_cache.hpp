// matrix_cache.hpp -- binary cache of a loaded matrix, next to the MTX text file.
//
// The reference re-parses the Matrix-Market text on every run of every host (readMtx, utils.hpp:474-520:
// fscanf per entry, minutes at 10^7 rows) and re-packs the BS-CSR packets in every SpMV constructor
// (host_spmv_bscsr.cpp:133-248).  A sweep (test_spmv_topk.py) runs the same matrix dozens of times, so the
// loaded CSR and, for the fixed-point engine, the packed partitions are kept in a versioned binary container:
//
//   header  (64 bytes)  magic "TKSMAT01" | kind (1 = CSR fp32, 2 = BS-CSR packets) | rows | cols | nnz |
//                       aux0 (CSR: tag of the source file, 0 = untagged; BS-CSR: partitions) | aux1 (BS-CSR: fixed_width) | payload bytes |
//                       FNV-1a 64 checksum of the payload
//   payload             CSR:    ptr[rows+1] u64 | idx[nnz] u32 | val[nnz] f32
//                       BS-CSR: packets_per_part[P] u64 | first_row[P] u32 | nnz_per_part[P] u64 | packets (64 B each)
// Little-endian, no alignment padding inside the payload.  Truncated or altered files are rejected.
#pragma once

#include <sys/stat.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace tkshost {

constexpr char kCacheMagic[8] = {'T', 'K', 'S', 'M', 'A', 'T', '0', '1'};
constexpr uint32_t kCacheCsr = 1, kCacheBscsr = 2;

struct CacheHeader {
    char magic[8];
    uint32_t kind, cols;
    uint64_t rows, nnz, aux0, aux1, payload_bytes, checksum;
};
static_assert(sizeof(CacheHeader) == 64, "cache header is 64 bytes");

inline uint64_t fnv1a64(const void *data, size_t n, uint64_t h = 1469598103934665603ull) {
    // 8 bytes per step (a byte-wise FNV over gigabytes would dominate the load time)
    const uint8_t *p = static_cast<const uint8_t *>(data);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h = (h ^ w) * 1099511628211ull;
    }
    for (; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

// What a CSR cache was made from: path, size and modification time of the Matrix-Market file and the two loader flags
// that change what the same text parses to (-z index base, -v ignore values).  A cache whose tag differs from the tag
// of the matrix a run asks for is stale or foreign and must not be used.  Never 0 (0 = "untagged").
inline uint64_t cache_source_tag(const char *source_path, int zero_indexed, int ignore_values);

struct CacheSection { const void *data; size_t bytes; };

inline int cache_write(const char *path, CacheHeader hdr, const std::vector<CacheSection> &sections, std::string *err) {
    std::memcpy(hdr.magic, kCacheMagic, 8);
    hdr.payload_bytes = 0;
    hdr.checksum = 1469598103934665603ull;
    for (const auto &s : sections) { hdr.payload_bytes += s.bytes; hdr.checksum = fnv1a64(s.data, s.bytes, hdr.checksum); }
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f) { if (err) *err = "cannot create " + tmp; return -1; }
    bool ok = std::fwrite(&hdr, sizeof hdr, 1, f) == 1;
    for (const auto &s : sections) ok = ok && (s.bytes == 0 || std::fwrite(s.data, 1, s.bytes, f) == s.bytes);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok || std::rename(tmp.c_str(), path) != 0) { std::remove(tmp.c_str()); if (err) *err = "short write to " + std::string(path); return -1; }
    return 0;
}

// Reads and validates the header; the payload follows at offset 64.
inline int cache_open(const char *path, uint32_t kind, CacheHeader *hdr, FILE **out, std::string *err) {
    FILE *f = std::fopen(path, "rb");
    if (!f) { if (err) *err = std::string("File ") + path + " not found"; return -1; }
    auto fail = [&](const char *m) { std::fclose(f); if (err) *err = std::string(path) + ": " + m; return -1; };
    if (std::fread(hdr, sizeof *hdr, 1, f) != 1) return fail("truncated header");
    if (std::memcmp(hdr->magic, kCacheMagic, 8) != 0) return fail("not a TKSMAT01 matrix cache");
    if (hdr->kind != kind) return fail("cache holds a different kind of matrix");
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    if (sz < 0 || (uint64_t)sz != sizeof *hdr + hdr->payload_bytes) return fail("truncated or oversized payload");
    std::fseek(f, (long)sizeof *hdr, SEEK_SET);
    *out = f;
    return 0;
}

inline int cache_read_sections(FILE *f, const CacheHeader &hdr, const std::vector<std::pair<void *, size_t>> &sections,
                               const char *path, std::string *err) {
    uint64_t h = 1469598103934665603ull, total = 0;
    for (const auto &s : sections) {
        if (s.second && std::fread(s.first, 1, s.second, f) != s.second) { std::fclose(f); if (err) *err = std::string(path) + ": short read"; return -1; }
        h = fnv1a64(s.first, s.second, h);
        total += s.second;
    }
    std::fclose(f);
    if (total != hdr.payload_bytes) { if (err) *err = std::string(path) + ": section sizes do not add up to the payload"; return -1; }
    if (h != hdr.checksum) { if (err) *err = std::string(path) + ": checksum mismatch (corrupted cache)"; return -1; }
    return 0;
}

inline uint64_t cache_source_tag(const char *source_path, int zero_indexed, int ignore_values) {
    struct stat st {};
    uint64_t meta[4] = {0, 0, (uint64_t)(zero_indexed != 0), (uint64_t)(ignore_values != 0)};
    if (::stat(source_path, &st) == 0) {
        meta[0] = (uint64_t)st.st_size;
        meta[1] = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    }
    uint64_t h = fnv1a64(source_path, std::strlen(source_path));
    h = fnv1a64(meta, sizeof meta, h);
    return h ? h : 1;
}

}  // namespace tkshost
