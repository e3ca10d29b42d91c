// spmv.hpp -- `struct SpMV`, the accelerator functor every reference main() drives, re-seated on the
// C ABI of libtopkspmv.so (include/topkspmv.h).  Same four verbs, same argument meaning:
//   ctor          upload the matrix (and the first query)      host_spmv_bscsr.cpp:104, host_spmv_topk_csr_gpu.cu:95
//   operator()    run, return elapsed nanoseconds of the kernel host_spmv_bscsr.cpp:323, host_spmv_topk_csr_gpu.cu:171
//   read_result   sorted (score desc) values and indices        host_spmv_bscsr.cpp:399, host_spmv_topk_csr_gpu.cu:233
//   reset         upload a new query, return elapsed ns         host_spmv_bscsr.cpp:450, host_spmv_topk_csr_gpu.cu:241
// Errors: the reference exits; so does this wrapper (message from tks_last_error).
#pragma once

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/topkspmv.h"
#include "bscsr_packer.hpp"
#include "fixed_point.hpp"
#include "types.hpp"

#define TKS_OR_DIE(h, call)                                                             \
    do {                                                                                \
        int rc__ = (call);                                                              \
        if (rc__ != 0) {                                                                \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, tks_last_error(h));    \
            exit(EXIT_FAILURE);                                                         \
        }                                                                               \
    } while (0)

// Exact fp32 engine: constructor signature of the reference GPU host (CSR arrays + query + k).
struct SpMV {
    tks_handle *h = nullptr;
    int_type num_rows, num_cols, num_nnz;
    int k;
    float last_full_ms = 0.f;

    SpMV(int_type *ptr, int_type *idx, float *val, int_type num_rows_, int_type num_cols_, int_type num_nnz_,
         float *vec, int k_, int device = 0, bool tie_higher = false, int debug = 0,
         bool use_half_precision_gpu = false)   // the reference ctor's last argument (host_spmv_topk_csr_gpu.cu:95)
        : num_rows(num_rows_), num_cols(num_cols_), num_nnz(num_nnz_), k(k_) {
        tks_config cfg;
        tks_default_config(&cfg);
        cfg.mode = TKS_MODE_FLOAT_CSR;
        cfg.device = device;
        cfg.max_cols = num_cols_ > MAX_COLS ? (int)num_cols_ : MAX_COLS;
        cfg.tie_break = tie_higher ? TKS_TIE_HIGHER_INDEX : TKS_TIE_LOWER_INDEX;
        cfg.value_type = use_half_precision_gpu ? TKS_VALUE_FP16 : TKS_VALUE_FP32;
        TKS_OR_DIE(nullptr, tks_create(&cfg, &h));
        if (debug) printf("Write inputs into device memory\n");
        TKS_OR_DIE(h, tks_upload_csr(h, num_rows, num_cols, num_nnz, ptr, 32, idx, val, 0));
        TKS_OR_DIE(h, tks_set_query(h, vec, 1));
    }
    ~SpMV() { tks_destroy(h); }
    SpMV(const SpMV &) = delete;

    float operator()(int debug) {
        float kernel_ms = 0.f;
        TKS_OR_DIE(h, tks_run(h, (uint32_t)k, &kernel_ms, &last_full_ms));
        if (debug) printf("Kernel terminated\nComputation took %f ms (%f ms with read-back)\n", kernel_ms, last_full_ms);
        return kernel_ms * 1e6f;
    }
    void read_result(std::vector<float> &res, std::vector<int_type> &res_idx, int debug = 0) {
        (void)debug;
        res.resize(k); res_idx.resize(k);
        uint32_t count = 0;
        TKS_OR_DIE(h, tks_read_result(h, 0, res_idx.data(), res.data(), &count));
    }
    long reset(float *vec, int debug) {
        auto t0 = std::chrono::high_resolution_clock::now();
        TKS_OR_DIE(h, tks_set_query(h, vec, 1));
        long ns = (long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::high_resolution_clock::now() - t0).count();
        if (debug) printf("Reset took %f ms\n", ns / 1e6);
        return ns;
    }
    // Throughput form (not in the reference, whose hosts keep one query in flight): reset + operator() without waiting,
    // read_result by ticket.  Up to kInFlight queries are kept; fetch(t) before submit number t + kInFlight.
    static constexpr int kInFlight = 3;
    uint64_t submit(float *vec) {
        uint64_t ticket = 0;
        TKS_OR_DIE(h, tks_submit_host(h, vec, (uint32_t)k, 0, &ticket));
        return ticket;
    }
    void fetch(uint64_t ticket, std::vector<float> &res, std::vector<int_type> &res_idx) {
        res.resize(k); res_idx.resize(k);
        uint32_t count = 0;
        TKS_OR_DIE(h, tks_fetch(h, ticket, res_idx.data(), res.data(), &count));
    }
};

// The same functor over SEVERAL GPUs driven by this one process (tks_group_*): rows sharded contiguously over the
// devices, K candidates exchanged over NVLink by the select kernels; the four verbs are unchanged.
struct SpMVGroup {
    tks_group *g = nullptr;
    int_type num_rows, num_cols, num_nnz;
    int k;
    float last_full_ms = 0.f;

#define TKS_G_OR_DIE(call)                                                                       \
    do {                                                                                         \
        int rc__ = (call);                                                                       \
        if (rc__ != 0) {                                                                         \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, tks_group_last_error(g));       \
            exit(EXIT_FAILURE);                                                                  \
        }                                                                                        \
    } while (0)

    SpMVGroup(int_type *ptr, int_type *idx, float *val, int_type num_rows_, int_type num_cols_, int_type num_nnz_,
              float *vec, int k_, const std::vector<int> &devices, bool tie_higher = false, int debug = 0,
              bool use_half_precision_gpu = false)
        : num_rows(num_rows_), num_cols(num_cols_), num_nnz(num_nnz_), k(k_) {
        tks_config cfg;
        tks_default_config(&cfg);
        cfg.mode = TKS_MODE_FLOAT_CSR;
        cfg.max_cols = num_cols_ > MAX_COLS ? (int)num_cols_ : MAX_COLS;
        cfg.tie_break = tie_higher ? TKS_TIE_HIGHER_INDEX : TKS_TIE_LOWER_INDEX;
        cfg.value_type = use_half_precision_gpu ? TKS_VALUE_FP16 : TKS_VALUE_FP32;
        std::vector<int32_t> dev(devices.begin(), devices.end());
        TKS_G_OR_DIE(tks_group_create(&cfg, dev.data(), (uint32_t)dev.size(), &g));
        if (debug) printf("Write inputs into the memory of %zu devices\n", dev.size());
        TKS_G_OR_DIE(tks_group_upload_csr(g, num_rows, num_cols, num_nnz, ptr, 32, idx, val));
        TKS_G_OR_DIE(tks_group_set_query(g, vec));
    }
    ~SpMVGroup() { tks_group_destroy(g); }
    SpMVGroup(const SpMVGroup &) = delete;

    float operator()(int debug) {
        float kernel_ms = 0.f;
        TKS_G_OR_DIE(tks_group_run(g, (uint32_t)k, &kernel_ms, &last_full_ms));
        if (debug) printf("Kernels terminated\nComputation took %f ms on the slowest device (%f ms with read-back)\n", kernel_ms, last_full_ms);
        return kernel_ms * 1e6f;
    }
    void read_result(std::vector<float> &res, std::vector<int_type> &res_idx, int debug = 0) {
        (void)debug;
        res.resize(k); res_idx.resize(k);
        uint32_t count = 0;
        TKS_G_OR_DIE(tks_group_read_result(g, 0, res_idx.data(), res.data(), &count));
    }
    long reset(float *vec, int debug) {
        auto t0 = std::chrono::high_resolution_clock::now();
        TKS_G_OR_DIE(tks_group_set_query(g, vec));
        long ns = (long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::high_resolution_clock::now() - t0).count();
        if (debug) printf("Reset took %f ms\n", ns / 1e6);
        return ns;
    }
    static constexpr int kInFlight = 3;
    uint64_t submit(float *vec) {
        uint64_t ticket = 0;
        TKS_G_OR_DIE(tks_group_submit_host(g, vec, (uint32_t)k, &ticket));
        return ticket;
    }
    void fetch(uint64_t ticket, std::vector<float> &res, std::vector<int_type> &res_idx) {
        res.resize(k); res_idx.resize(k);
        uint32_t count = 0;
        TKS_G_OR_DIE(tks_group_fetch(g, ticket, res_idx.data(), res.data(), &count));
    }
#undef TKS_G_OR_DIE
};

// FPGA-semantics engine: constructor signature of the reference FPGA host (COO + fixed-point values).
struct SpMVFixed {
    tks_handle *h = nullptr;
    int_type num_rows, num_cols, num_nnz;
    int k;
    int fixed_width, partitions;
    float last_full_ms = 0.f;

    SpMVFixed(int_type *x, int_type *y, ufixed32 *val, int_type num_rows_, int_type num_cols_, int_type num_nnz_,
              ufixed32 *vec, int k_, int fixed_width_ = FIXED_WIDTH, int partitions_ = SPMV_PARTITIONS,
              int local_k = K, int lfr = LIMITED_FINISHED_ROWS, int device = 0, int debug = 0, bool drift_free = false,
              bool device_pack = false)
        : num_rows(num_rows_), num_cols(num_cols_), num_nnz(num_nnz_), k(k_), fixed_width(fixed_width_), partitions(partitions_) {
        tks_config cfg;
        tks_default_config(&cfg);
        cfg.mode = TKS_MODE_FIXED_BSCSR;
        cfg.fixed_width = fixed_width;
        cfg.partitions = partitions;
        cfg.local_k = local_k;
        cfg.limited_finished_rows = lfr;
        cfg.device = device;
        cfg.tie_break = TKS_TIE_HIGHER_INDEX;   // sort_tuples order of the reference merge (host:447)
        cfg.fixed_drift_free = drift_free ? 1 : 0;
        TKS_OR_DIE(nullptr, tks_create(&cfg, &h));
        static_assert(sizeof(ufixed32) == 4, "ufixed32 must be a bare 32-bit word");
        if (device_pack) {
            // packet_coo (host:133-187) and the device tables, built on the GPU from the COO arrays
            if (debug) printf("Pack and write inputs on the device\n");
            TKS_OR_DIE(h, tks_upload_coo_fixed(h, x, y, reinterpret_cast<uint32_t *>(val), num_nnz, num_rows, num_cols));
            TKS_OR_DIE(h, tks_set_query(h, vec, 1));
            return;
        }
        // packet_coo (host:133-187)
        std::vector<uint64_t> ppp(partitions), npp(partitions);
        std::vector<uint32_t> first_row(partitions);
        TKS_OR_DIE(h, tks_pack_bscsr(x, y, reinterpret_cast<uint32_t *>(val), num_nnz, num_rows, partitions, fixed_width,
                                     ppp.data(), first_row.data(), npp.data(), nullptr));
        uint64_t total = 0;
        for (auto n : ppp) total += n;
        std::vector<tkshost::Packet512> packets(total);
        TKS_OR_DIE(h, tks_pack_bscsr(x, y, reinterpret_cast<uint32_t *>(val), num_nnz, num_rows, partitions, fixed_width,
                                     ppp.data(), first_row.data(), npp.data(), packets.data()));
        std::vector<const void *> pp(partitions);
        uint64_t off = 0;
        for (int p = 0; p < partitions; p++) { pp[p] = packets.data() + off; off += ppp[p]; }
        if (debug) printf("Write inputs into device memory (%llu packets)\n", (unsigned long long)total);
        TKS_OR_DIE(h, tks_upload_bscsr(h, num_cols, (uint32_t)partitions, ppp.data(), pp.data(), first_row.data(), npp.data()));
        TKS_OR_DIE(h, tks_set_query(h, vec, 1));
    }
    ~SpMVFixed() { tks_destroy(h); }
    SpMVFixed(const SpMVFixed &) = delete;

    long operator()(int debug) {
        float kernel_ms = 0.f;
        TKS_OR_DIE(h, tks_run(h, (uint32_t)k, &kernel_ms, &last_full_ms));
        if (debug) printf("Kernel terminated\nComputation took %f ms\n", kernel_ms);
        return (long)(kernel_ms * 1e6f);
    }
    // appends, and may return fewer than k entries, like host_spmv_bscsr.cpp:399-448
    void read_result(std::vector<ufixed32> &res, std::vector<int_type> &res_idx, int debug = 0) {
        (void)debug;
        std::vector<uint32_t> idx(k), val(k);
        uint32_t count = 0;
        TKS_OR_DIE(h, tks_read_result(h, 0, idx.data(), val.data(), &count));
        for (uint32_t i = 0; i < count; i++) { res_idx.push_back(idx[i]); res.push_back(ufixed32::from_raw(val[i])); }
    }
    long reset(ufixed32 *vec, int debug) {
        auto t0 = std::chrono::high_resolution_clock::now();
        TKS_OR_DIE(h, tks_set_query(h, vec, 1));
        long ns = (long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::high_resolution_clock::now() - t0).count();
        if (debug) printf("Reset took %f ms\n", ns / 1e6);
        return ns;
    }
    // Throughput form: two queries are kept in this mode (the merge of a query runs in fetch while the next one streams)
    static constexpr int kInFlight = 1;
    uint64_t submit(ufixed32 *vec) {
        uint64_t ticket = 0;
        TKS_OR_DIE(h, tks_submit_host(h, vec, (uint32_t)k, 0, &ticket));
        return ticket;
    }
    void fetch(uint64_t ticket, std::vector<ufixed32> &res, std::vector<int_type> &res_idx) {
        std::vector<uint32_t> idx(k), val(k);
        uint32_t count = 0;
        TKS_OR_DIE(h, tks_fetch(h, ticket, idx.data(), val.data(), &count));
        res.clear(); res_idx.clear();
        for (uint32_t i = 0; i < count; i++) { res_idx.push_back(idx[i]); res.push_back(ufixed32::from_raw(val[i])); }
    }
};
