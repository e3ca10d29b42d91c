// ref_gold.cpp -- thin extern "C" driver around the REFERENCE's own headers.
//
// TEST INFRASTRUCTURE ONLY.  Compiled from the sources where they lie under
// /root/reference (see oracle/Makefile, target `ref`); output goes to
// oracle/_ref/libref_gold.so (git-ignored, travels to the GPU box).  Nothing
// of the reference is copied: this file only #includes it, exactly the way
// src/gpu/host_spmv_topk_csr_gpu.cu:15-19,28-29 does (real_type = float).
//
// It pins oracle/oracle.c's float path and the host C++ mirror
// (MTX loader, coo2csr, sort_tuples, create_sample_vector).
#include <cstdint>
#include <cstring>
#include <vector>

#include "src/common/utils/utils.hpp"
#include "src/common/utils/options.hpp"
#include "src/common/utils/evaluation_utils.hpp"
#include "src/fpga/src/ip/coo_matrix.hpp"
#include "src/fpga/src/gold_algorithms/gold_algorithms.hpp"

#define REF_API extern "C" __attribute__((visibility("default")))

// gold_algorithms.hpp:188-246 + evaluation_utils.hpp:40-62, as sw_test() calls them
// (host_spmv_topk_csr_gpu.cu:268-285).
REF_API void ref_gold_topk_f32(const uint32_t *row, const uint32_t *col, const float *val,
                               uint64_t nnz, const float *vec, int k, int sort,
                               uint32_t *res_idx, float *res_val) {
    std::vector<unsigned> x(row, row + nnz), y(col, col + nnz);
    std::vector<float> v(val, val + nnz);
    coo_t<unsigned, float> coo(x, y, v);
    spmv_coo_gold_top_k(coo, const_cast<float *>(vec), k, res_idx, res_val);
    if (sort) sort_tuples((size_t)k, res_idx, res_val);
}

REF_API void ref_sort_tuples_f32(uint32_t n, uint32_t *idx, float *val) { sort_tuples((size_t)n, idx, val); }

// utils.hpp:474-520 readMtx (+ mmio.hpp).  Two-call protocol: first with null
// outputs to get nnz, then with buffers.  Returns 0 on success.
static std::vector<unsigned> g_x, g_y;
static std::vector<float> g_v;
REF_API int ref_read_mtx(const char *path, int zero_indexed, int sort, uint32_t *rows, uint32_t *cols,
                         uint64_t *nnz) {
    g_x.clear(); g_y.clear(); g_v.clear();
    unsigned r = 0, c = 0, n = 0;
    int rc = readMtx<unsigned, float>(path, &g_x, &g_y, &g_v, &r, &c, &n, 0, true, false, zero_indexed != 0,
                                      sort != 0);
    *rows = r; *cols = c; *nnz = g_x.size();
    return rc;
}
REF_API void ref_read_mtx_fetch(uint32_t *x, uint32_t *y, float *v) {
    std::memcpy(x, g_x.data(), g_x.size() * 4);
    std::memcpy(y, g_y.data(), g_y.size() * 4);
    std::memcpy(v, g_v.data(), g_v.size() * 4);
}

// utils.hpp:522-580 coo2csr (sort_tuples=false, as host_spmv_topk_csr_gpu.cu:337 calls it).
REF_API void ref_coo2csr(const uint32_t *row, const uint32_t *col, const float *val, uint64_t nnz,
                         uint32_t nrows, uint32_t ncols, uint32_t *ptr, uint32_t *idx, float *out_val) {
    std::vector<unsigned> x(row, row + nnz), y(col, col + nnz);
    std::vector<float> v(val, val + nnz);
    coo2csr<unsigned, float>(ptr, idx, out_val, x, y, v, nrows, ncols, false);
}

// utils.hpp:234-267 create_sample_vector<float>(vec, size, random=true, sum_to_one=false, norm_one=true, seed).
REF_API void ref_create_sample_vector_f32(float *vec, int size, int seed) {
    create_sample_vector<float>(vec, size, true, false, true, seed);
}

// coo_matrix.hpp:21-27 num_rows = max(start)+1.
REF_API uint32_t ref_coo_num_rows(const uint32_t *row, uint64_t nnz) {
    std::vector<unsigned> x(row, row + nnz), y(nnz, 0u);
    std::vector<float> v(nnz, 0.f);
    coo_t<unsigned, float> coo(x, y, v);
    return coo.num_rows;
}

// evaluation_utils.hpp:273-297 mean / st_dev with skip.
REF_API float ref_mean(const float *x, int n, int skip) { return mean(std::vector<float>(x, x + n), skip); }
REF_API float ref_st_dev(const float *x, int n, int skip) { return st_dev(std::vector<float>(x, x + n), skip); }
