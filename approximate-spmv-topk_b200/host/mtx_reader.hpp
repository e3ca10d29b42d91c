// mtx_reader.hpp -- Matrix-Market loader of the host surface.
//
// Mirrors readMtx(fname, &rows, &cols, &vals, &nrows, &ncols, &nnz, directed, read_values, debug,
// zero_indexed_file, sort_tuples) of the reference (src/common/utils/utils.hpp:474-520, with
// readTuples :372-404, customSort :350-370 and the NIST banner parser src/common/utils/mmio.hpp:124-230):
// coordinate format, real / integer / pattern fields, general or symmetric storage, optional -1 index
// shift, optional (row, col) sort.  Differences, all deliberate:
//   * the index base is an explicit argument and defaults to the MTX standard (1-based); the reference
//     hosts hard-code zero_indexed_file=true (host_spmv_bscsr.cpp:539) although its own generator
//     writes 1-based files (create_matrices.py:120,124) -- SURVEY section 4, sharp edge 1;
//   * errors are returned, not exit(1);
//   * the body is parsed from one in-memory buffer by several threads instead of fscanf per token
//     (minutes -> seconds at 2e8 entries).
// Written from the format specification; no code is shared with mmio.hpp.
#pragma once

#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

namespace tkshost {

struct MtxInfo {
    bool coordinate = false, pattern = false, integer = false, real = false, symmetric = false;
    uint64_t rows = 0, cols = 0, entries = 0;
};

inline std::string lower(std::string s) {
    for (auto &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

// Returns 0 on success, negative on error (message in *err).
template <typename I, typename T>
inline int readMtx(const char *fname, std::vector<I> *row_indices, std::vector<I> *col_indices,
                   std::vector<T> *values, I *num_rows, I *num_cols, I *num_nnz, int directed = 1,
                   bool read_values = true, bool debug = false, bool zero_indexed_file = false,
                   bool sort_tuples = true, std::string *err = nullptr, unsigned num_threads = 0) {
    auto fail = [&](const std::string &m) {
        if (err) *err = m;
        if (debug) fprintf(stderr, "readMtx: %s\n", m.c_str());
        return -1;
    };
    FILE *f = fopen(fname, "rb");
    if (!f) return fail(std::string("File ") + fname + " not found");
    fseek(f, 0, SEEK_END);
    long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)fsz + 1);
    if (fsz > 0 && fread(buf.data(), 1, (size_t)fsz, f) != (size_t)fsz) { fclose(f); return fail("short read"); }
    fclose(f);
    buf[(size_t)fsz] = '\n';
    const char *p = buf.data(), *end = buf.data() + fsz;

    // banner: %%MatrixMarket matrix coordinate <field> <symmetry>
    MtxInfo info;
    {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!eol) return fail("Could not process Matrix Market banner.");
        std::string line(p, eol);
        char a[64] = {0}, b[64] = {0}, c[64] = {0}, d[64] = {0}, e[64] = {0};
        if (sscanf(line.c_str(), "%63s %63s %63s %63s %63s", a, b, c, d, e) != 5 || strcmp(a, "%%MatrixMarket") != 0)
            return fail("Could not process Matrix Market banner.");
        if (lower(b) != "matrix") return fail("unsupported MatrixMarket object");
        info.coordinate = lower(c) == "coordinate";
        std::string fld = lower(d), sym = lower(e);
        info.pattern = fld == "pattern"; info.integer = fld == "integer"; info.real = fld == "real";
        info.symmetric = sym == "symmetric";
        if (!info.coordinate) return fail("only coordinate (sparse) MatrixMarket files are supported");
        if (!(info.pattern || info.integer || info.real)) return fail("unsupported MatrixMarket field " + fld);
        if (!(info.symmetric || sym == "general")) return fail("unsupported MatrixMarket symmetry " + sym);
        p = eol + 1;
    }
    // comments / blank lines, then the size line
    for (;;) {
        if (p >= end) return fail("premature end of file before size line");
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p + 1));
        if (*p == '%' || eol == p || (eol == p + 1 && *p == '\r')) { p = eol + 1; continue; }
        unsigned long long r, c, n;
        if (sscanf(std::string(p, eol).c_str(), "%llu %llu %llu", &r, &c, &n) != 3) return fail("bad size line");
        info.rows = r; info.cols = c; info.entries = n;
        p = eol + 1;
        break;
    }

    const bool has_value = !info.pattern;
    unsigned T_ = num_threads ? num_threads : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if ((uint64_t)(end - p) < (1u << 20)) T_ = 1;
    std::vector<const char *> cut(T_ + 1);
    cut[0] = p; cut[T_] = end;
    for (unsigned t = 1; t < T_; t++) {
        const char *q = p + (size_t)(end - p) * t / T_;
        const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q + 1));
        cut[t] = (nl && nl < end) ? nl + 1 : end;
    }
    struct Part { std::vector<I> r, c; std::vector<T> v; bool bad = false; };
    std::vector<Part> parts(T_);
    auto work = [&](unsigned t) {
        Part &o = parts[t];
        const char *q = cut[t], *qe = cut[t + 1];
        size_t guess = (size_t)(qe - q) / 16 + 16;
        o.r.reserve(guess); o.c.reserve(guess); o.v.reserve(guess);
        auto skip_ws = [&]() { while (q < qe && (*q == ' ' || *q == '\t' || *q == '\r' || *q == '\n')) q++; };
        while (true) {
            skip_ws();
            if (q >= qe) break;
            if (*q == '%') { while (q < qe && *q != '\n') q++; continue; }
            unsigned long long ri = 0, ci = 0;
            auto r1 = std::from_chars(q, qe, ri);
            if (r1.ec != std::errc()) { o.bad = true; return; }
            q = r1.ptr; while (q < qe && (*q == ' ' || *q == '\t')) q++;
            auto r2 = std::from_chars(q, qe, ci);
            if (r2.ec != std::errc()) { o.bad = true; return; }
            q = r2.ptr;
            double val = 1.0;
            if (has_value) {
                while (q < qe && (*q == ' ' || *q == '\t')) q++;
                if (q < qe && *q == '+') q++;
                auto r3 = std::from_chars(q, qe, val);
                if (r3.ec != std::errc()) { o.bad = true; return; }
                q = r3.ptr;
            }
            while (q < qe && *q != '\n') q++;
            if (!zero_indexed_file) { ri--; ci--; }          // utils.hpp:396-399
            o.r.push_back((I)ri); o.c.push_back((I)ci);
            o.v.push_back(read_values ? (T)val : (T)1.0);   // utils.hpp:387-391,401
        }
    };
    if (T_ == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T_; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    size_t total = 0;
    for (auto &o : parts) { if (o.bad) return fail("malformed entry line"); total += o.r.size(); }
    if (total < info.entries) return fail("Error: Not enough rows in mtx file!");
    total = (size_t)info.entries;   // like the reference, read exactly nnz entries
    row_indices->clear(); col_indices->clear(); values->clear();
    row_indices->reserve(total); col_indices->reserve(total); values->reserve(total);
    for (auto &o : parts) {
        size_t take = std::min(o.r.size(), total - row_indices->size());
        row_indices->insert(row_indices->end(), o.r.begin(), o.r.begin() + take);
        col_indices->insert(col_indices->end(), o.c.begin(), o.c.begin() + take);
        values->insert(values->end(), o.v.begin(), o.v.begin() + take);
    }
    // symmetric expansion (utils.hpp:504-509 with undirect :406-419)
    bool is_undirected = info.symmetric || directed == 2;
    if (directed == 1) is_undirected = false;
    if (is_undirected) {
        size_t n0 = row_indices->size();
        for (size_t i = 0; i < n0; i++) {
            if ((*col_indices)[i] != (*row_indices)[i]) {
                row_indices->push_back((*col_indices)[i]);
                col_indices->push_back((*row_indices)[i]);
                values->push_back((*values)[i]);
            }
        }
    }
    if (sort_tuples) {   // customSort: by (row, col)
        std::vector<size_t> perm(row_indices->size());
        for (size_t i = 0; i < perm.size(); i++) perm[i] = i;
        std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) {
            if ((*row_indices)[a] != (*row_indices)[b]) return (*row_indices)[a] < (*row_indices)[b];
            return (*col_indices)[a] < (*col_indices)[b];
        });
        std::vector<I> r2(perm.size()), c2(perm.size());
        std::vector<T> v2(perm.size());
        for (size_t i = 0; i < perm.size(); i++) { r2[i] = (*row_indices)[perm[i]]; c2[i] = (*col_indices)[perm[i]]; v2[i] = (*values)[perm[i]]; }
        row_indices->swap(r2); col_indices->swap(c2); values->swap(v2);
    }
    *num_rows = (I)info.rows;
    *num_cols = (I)info.cols;
    *num_nnz = (I)row_indices->size();
    return 0;
}

// coo2csr (utils.hpp:522-580): counting sort by row, stable inside a row.
template <typename I, typename T>
inline int coo2csr(I *csrRowPtr, I *csrColInd, T *csrVal, const std::vector<I> &row_indices,
                   const std::vector<I> &col_indices, const std::vector<T> &values, I nrows, I ncols) {
    const size_t nvals = row_indices.size();
    for (size_t i = 0; i <= (size_t)nrows; i++) csrRowPtr[i] = 0;
    for (size_t i = 0; i < nvals; i++) {
        if (row_indices[i] >= nrows || col_indices[i] >= ncols) return -1;
        csrRowPtr[(size_t)row_indices[i] + 1]++;
    }
    for (size_t r = 0; r < (size_t)nrows; r++) csrRowPtr[r + 1] += csrRowPtr[r];
    std::vector<I> next(csrRowPtr, csrRowPtr + nrows);
    for (size_t i = 0; i < nvals; i++) {
        I dst = next[row_indices[i]]++;
        csrColInd[dst] = col_indices[i];
        csrVal[dst] = values[i];
    }
    return 0;
}

}  // namespace tkshost
