// ref_fpga.cpp -- the REFERENCE's FPGA host and HLS kernel, compiled as they lie and run in software.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile (target `ref`) into
// oracle/_ref/libref_fpga_w<W>_k<K>_l<LFR>_p<P>.so, one library per knob combination, because the
// reference's knobs are compile-time macros (src/common/types.hpp:20,36,51,77).  Nothing of the
// reference is copied: this file #includes
//     src/fpga/src/host_spmv_bscsr.cpp                      (struct SpMV: partitioning, packet builder,
//                                                            kernel launch, read_result merge; main renamed)
//     src/fpga/src/ip/spmv/spmv_bscsr_top_k_multicore.cpp   (the HLS kernel, spmv_bscsr_top_k_main)
// against the header stand-ins in oracle/shim/ (ap_int.h, ap_fixed.h, hls_stream.h, CL/cl2.hpp): the Vitis
// and OpenCL headers those sources need are absent from this image.  The shim runs the kernel the way
// Vitis sw_emu does -- cl::CommandQueue::enqueueTask calls the kernel's C function.
//
// Knob override: types.hpp is `#pragma once` and every derived macro (SCALE, BSCSR_PACKET_SIZE,
// PADDING_SIZE, VEC_REPLICAS, SUPER_SPMV_PARTITIONS ...) expands lazily, so including it first and
// re-defining the four primary macros re-parameterises every later use without touching the file.
//
// It pins oracle/oracle.c's fixed-point path (quantisation, packet builder, kernel, merge) and provides
// the fixtures under tests/golden/.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "src/common/types.hpp"
#ifdef REF_W
#undef FIXED_WIDTH
#define FIXED_WIDTH REF_W
#endif
#ifdef REF_K
#undef K
#define K REF_K
#endif
#ifdef REF_LFR
#undef LIMITED_FINISHED_ROWS
#define LIMITED_FINISHED_ROWS REF_LFR
#endif
#ifdef REF_P
#undef SPMV_PARTITIONS
#define SPMV_PARTITIONS REF_P
#endif

#define main ref_fpga_unused_main
#include "src/fpga/src/host_spmv_bscsr.cpp"
#undef main
#include "src/fpga/src/ip/spmv/spmv_bscsr_top_k_multicore.cpp"

// enqueueTask -> the kernel's C function, arguments in the order SpMV::setup binds them
// (host_spmv_bscsr.cpp:279-310 <-> spmv_bscsr_top_k_multicore.cpp:8-34).
void apshim_cl_task(const std::vector<cl::KernelArg> &a) {
    spmv_bscsr_top_k_main(
        (input_block *)a[0].ptr, (input_block *)a[1].ptr, (input_block *)a[2].ptr, (input_block *)a[3].ptr,
        (int_type)a[4].scalar, (int_type)a[5].scalar, (int_type)a[6].scalar, (int_type)a[7].scalar,
        (int_type)a[8].scalar, (int_type)a[9].scalar, (int_type)a[10].scalar, (int_type)a[11].scalar,
        (int_type)a[12].scalar, (int_type)a[13].scalar, (int_type)a[14].scalar, (int_type)a[15].scalar,
        (vec_real_inout_bscsr *)a[16].ptr,
        (input_packet_int_bscsr *)a[17].ptr, (input_packet_int_bscsr *)a[18].ptr, (input_packet_int_bscsr *)a[19].ptr,
        (input_packet_int_bscsr *)a[20].ptr,
        (input_packet_real_inout_bscsr *)a[21].ptr, (input_packet_real_inout_bscsr *)a[22].ptr,
        (input_packet_real_inout_bscsr *)a[23].ptr, (input_packet_real_inout_bscsr *)a[24].ptr);
}

#define REF_API extern "C" __attribute__((visibility("default")))

static_assert(sizeof(real_type_inout) == 4, "ap_ufixed<32,1> must be one 32-bit word");
static_assert(sizeof(input_block) == 64, "input_block must be 64 bytes");
static_assert(sizeof(input_packet_int_bscsr) == 64 && sizeof(input_packet_real_inout_bscsr) == 64, "result packets are 64 bytes");

static inline real_type_inout from_raw32(uint32_t r) { real_type_inout t; t.V = r; return t; }

struct RefFpga {
    std::vector<int_type> x, y;
    std::vector<real_type_inout> val, vec;
    ConfigOpenCL config;
    SpMV *spmv = nullptr;
    RefFpga() : config("spmv_bscsr_top_k_main", SUPER_SPMV_PARTITIONS, OPENCL_QUEUES) {}
};

// knobs this library was compiled with
REF_API void ref_fpga_params(int *width, int *packet_size, int *local_k, int *lfr, int *partitions) {   // NB: gold_algorithms.hpp:274 #defines B
    *width = FIXED_WIDTH; *packet_size = BSCSR_PACKET_SIZE; *local_k = K; *lfr = LIMITED_FINISHED_ROWS; *partitions = SPMV_PARTITIONS;
}

// (real_type_inout) double -- readTuples, utils.hpp:401; create_sample_vector, utils.hpp:242
REF_API uint32_t ref_fx32_from_double(double v) { return (uint32_t)((real_type_inout)v).V; }
// write_block_val's cast (real_type) x.to_float() -- fpga_utils.hpp:336-338
REF_API uint32_t ref_fxW_from_fx32(uint32_t raw32) { real_type t = (real_type)from_raw32(raw32).to_float(); return (uint32_t)t.V; }
// create_sample_vector<real_type_inout>(vec, size, random, !sum_to_one, norm_one, seed) -- host_spmv_bscsr.cpp:547
REF_API void ref_create_sample_vector_fx32(uint32_t *out, int size, int seed) {
    std::vector<real_type_inout> v(size);
    create_sample_vector(v.data(), size, true, false, true, seed);
    for (int i = 0; i < size; i++) out[i] = (uint32_t)v[i].V;
}

// SpMV ctor (host_spmv_bscsr.cpp:104-131): partition, packet_coo, setup.  val32/vec32 are raw ap_ufixed<32,1> words.
REF_API void *ref_fpga_create(const uint32_t *x, const uint32_t *y, const uint32_t *val32, uint64_t nnz, uint32_t rows,
                              uint32_t cols, const uint32_t *vec32) {
    RefFpga *r = new RefFpga();
    r->x.assign(x, x + nnz);
    r->y.assign(y, y + nnz);
    r->val.resize(nnz);
    for (uint64_t i = 0; i < nnz; i++) r->val[i] = from_raw32(val32[i]);
    r->vec.resize(cols);
    for (uint32_t i = 0; i < cols; i++) r->vec[i] = from_raw32(vec32[i]);
    r->spmv = new SpMV(r->config, r->x.data(), r->y.data(), r->val.data(), rows, cols, (int_type)nnz, r->vec.data(), 0);
    return r;
}
REF_API void ref_fpga_destroy(void *h) {
    RefFpga *r = (RefFpga *)h;
    if (!r) return;
    delete r->spmv;
    delete r;
}
// partition bookkeeping (host_spmv_bscsr.cpp:144-150)
REF_API void ref_fpga_partition_info(void *h, int p, uint32_t *first_row, uint32_t *last_row, uint32_t *nnz, uint32_t *nblocks) {
    SubSpMVPartition &s = ((RefFpga *)h)->spmv->get_partition(p);
    *first_row = s.first_row; *last_row = s.last_row; *nnz = s.num_nnz_partition; *nblocks = s.num_blocks_nnz;
}
// the 64-byte packets packet_coo_partition produced (host_spmv_bscsr.cpp:189-248)
REF_API void ref_fpga_packets(void *h, int p, void *out) {
    SubSpMVPartition &s = ((RefFpga *)h)->spmv->get_partition(p);
    std::memcpy(out, s.coo_in.data(), s.coo_in.size() * sizeof(input_block));
}
// the packed query (host_spmv_bscsr.cpp:173-186)
REF_API void ref_fpga_query_blocks(void *h, void *out) {
    SpMV *s = ((RefFpga *)h)->spmv;
    std::memcpy(out, s->vec_in, s->num_blocks_cols * sizeof(vec_real_inout_bscsr));
}
// operator() (host_spmv_bscsr.cpp:323-397): all compute units, through the shim's enqueueTask
REF_API void ref_fpga_run(void *h) { (*((RefFpga *)h)->spmv)(0); }
// reset (host_spmv_bscsr.cpp:450-484)
REF_API void ref_fpga_reset(void *h, const uint32_t *vec32) {
    RefFpga *r = (RefFpga *)h;
    for (size_t i = 0; i < r->vec.size(); i++) r->vec[i] = from_raw32(vec32[i]);
    r->spmv->reset(r->vec.data(), 0);
}
// raw kernel output of partition p: K words of 16 x u32 each (spmv_bscsr_top_k_multicore.cpp:151-185)
REF_API void ref_fpga_result_words(void *h, int p, uint32_t *idx_words, uint32_t *val_words) {
    SubSpMVPartition &s = ((RefFpga *)h)->spmv->get_partition(p);
    std::memcpy(idx_words, s.res_idx_out.data(), K * TOPK_RES_COPIES * 64);
    std::memcpy(val_words, s.res_out.data(), K * TOPK_RES_COPIES * 64);
}
// read_result (host_spmv_bscsr.cpp:399-448): merged, de-duplicated, sort_tuples order
REF_API uint32_t ref_fpga_read_result(void *h, uint32_t *idx_out, uint32_t *val_out, uint32_t cap) {
    std::vector<real_type_inout> res;
    std::vector<int_type> res_idx;
    ((RefFpga *)h)->spmv->read_result(res, res_idx, 0);
    const uint32_t n = (uint32_t)res.size();
    for (uint32_t i = 0; i < n && i < cap; i++) { idx_out[i] = res_idx[i]; val_out[i] = (uint32_t)res[i].V; }
    return n;
}
// sw_test's top-k half (host_spmv_bscsr.cpp:497-501): spmv_coo_gold_top_k<int_type, real_type_inout> + sort_tuples
REF_API void ref_gold_topk_fx32(const uint32_t *x, const uint32_t *y, const uint32_t *val32, uint64_t nnz, const uint32_t *vec32,
                                uint32_t cols, int k, uint32_t *res_idx, uint32_t *res_val) {
    std::vector<int_type> xv(x, x + nnz), yv(y, y + nnz);
    std::vector<real_type_inout> v(nnz), vec(cols), out(k);
    for (uint64_t i = 0; i < nnz; i++) v[i] = from_raw32(val32[i]);
    for (uint32_t i = 0; i < cols; i++) vec[i] = from_raw32(vec32[i]);
    coo_t<int_type, real_type_inout> coo(xv, yv, v);
    std::vector<int_type> idx(k);
    spmv_coo_gold_top_k(coo, vec.data(), k, idx.data(), out.data());
    sort_tuples((size_t)k, idx.data(), out.data());
    for (int i = 0; i < k; i++) { res_idx[i] = idx[i]; res_val[i] = (uint32_t)out[i].V; }
}
