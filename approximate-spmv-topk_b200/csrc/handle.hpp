// handle.hpp -- the object behind tks_handle (one matrix shard on one device).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/topkspmv.h"
#include "common.cuh"

namespace tks {

struct BscsrState;   // bscsr_api.cu
struct RunState;     // csr_topk.cuh

struct Handle {
    tks_config cfg{};
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm0 = nullptr, evm1 = nullptr, ev_query = nullptr;
    std::string err;

    // matrix (float CSR mode)
    bool have_matrix = false;
    uint64_t rows = 0, nnz = 0, row_offset = 0;
    uint32_t cols = 0;
    void *d_val = nullptr;              // fp32 values, or IEEE halves when cfg.value_type == TKS_VALUE_FP16
    uint16_t *d_col16 = nullptr;        // column * 4
    uint32_t *d_col12 = nullptr;        // the same packed to 12 bits (cols <= 1024): streamed by the single-query kernels
    uint32_t *d_rowbits = nullptr;      // row-start bitmap, one bit per non-zero (zeroed words, atomicOr at build)
    uint64_t *d_ptr64 = nullptr;        // kept for tks_download_csr (exact copy of row_ptr as u64)
    uint64_t *d_chunk_start = nullptr;
    uint32_t *d_chunk_ord = nullptr;
    uint32_t *d_row_map = nullptr;      // ordinal -> row, only when the matrix has empty rows
    uint32_t n_chunks = 0, chunk_nnz = 0;
    uint64_t device_bytes = 0;

    // query + scratch + results (float mode), sized by max_batch
    uint32_t batch = 0;                 // queries set by the last tks_set_query
    bool have_query = false;
    float *d_x = nullptr;
    RunState *d_state = nullptr;        // one per query slot
    uint64_t *d_pool = nullptr;
    uint64_t pool_cap = 0;
    uint32_t *d_sample_keys = nullptr;
    uint32_t n_sample_cap = 0;
    uint32_t kmax = 0;
    uint64_t *d_res_keys = nullptr;
    uint32_t *d_res_block = nullptr, *h_res_block = nullptr;   // [count][idx][val], device / pinned host
    size_t res_block_bytes = 0;
    uint32_t *d_res_idx = nullptr;
    float *d_res_val = nullptr;
    uint32_t *d_res_count = nullptr;
    uint32_t *h_res_idx = nullptr;      // pinned
    float *h_res_val = nullptr;         // pinned
    uint32_t *h_res_count = nullptr;    // pinned
    float *h_x = nullptr;               // pinned staging for tks_set_query
    uint32_t last_k = 0;
    bool have_result = false;
    // blocking tks_run, one query: sample -> main -> select captured once per k into a CUDA graph and replayed with one
    // launch (api.cu run_graph_*); dropped whenever the matrix changes
    cudaGraphExec_t run_graph = nullptr;
    uint32_t run_graph_k = 0;
    bool run_graph_failed = false;
    bool res_on_host = false;           // the last select wrote indices / scores / count to the pinned host block only
    bool last_run_pipelined = false;

    // batched mode (csr_batched.cuh), allocated when max_batch > 1
    float *d_xT = nullptr;              // [npass][max_cols+1][32] transposed query tables
    uint64_t *d_bpool = nullptr;        // [max_batch][bpool_cap] candidate keys
    uint32_t bpool_cap = 0;
    uint32_t *d_pass_counter = nullptr; // [ceil(max_batch/32)]
    uint32_t *d_bsample_keys = nullptr; // [max_batch][b_sample_cap]
    uint32_t b_sample_cap = 0;
    bool batched_ok = false;            // the kernels' shared memory fits for max_cols
    bool last_run_batched = false;
    bool overflow_check_pending = false; // an async batched run has not been checked for pool overflow yet

    // peer-memory candidate exchange (several GPUs of one box; csr_topk.cuh peer_exchange_merge_kernel)
    uint64_t *d_peer_window = nullptr;  // this rank's window of tagged key records, exported through CUDA IPC
    void *peer_mapped[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // peers' windows
    uint32_t peer_world = 0, peer_rank = 0, peer_seq = 0;
    bool peer_ready = false;

    // launch geometry of the main kernel, per CAP variant
    int main_grid[4] = {0, 0, 0, 0};
    int tma_grid = 0;                    // grid of the bulk-copy variant of the k <= 128 main kernel (TKS_TMA=1)
    int x_grid = 0;                                                // TKS_XCOPIES=1 variant (0: does not fit, not used)

    // pipelined submits (tks_submit; api.cu): per-slot scratch so that consecutive queries overlap.  The sample and
    // the select kernels run on two engine-owned streams, the main kernels on the caller's; hand-over by sequence
    // numbers in RunState.  At most pipe_slots (<= kPipeSlots) queries are in flight.
    static constexpr int kPipeSlots = 4;
    int pipe_slots = kPipeSlots;                                   // TKS_PIPE_SLOTS (2..4): A/B switch
    cudaStream_t pipe_sample_stream = nullptr, pipe_select_stream = nullptr;
    cudaEvent_t pipe_ev_done[kPipeSlots] = {};                     // recorded after the slot's select kernel
    cudaEvent_t pipe_ev_query = nullptr;
    bool pipe_busy[kPipeSlots] = {};
    RunState *d_pipe_state = nullptr;                              // [kPipeSlots]
    uint64_t *d_pipe_pool[kPipeSlots] = {};
    uint32_t *d_pipe_sample_keys = nullptr;
    uint64_t *d_pipe_stamps = nullptr, *h_pipe_stamps = nullptr;   // %globaltimer at the end of every select, ring of kPipeStamps
    static constexpr uint32_t kPipeStamps = 4096;
    float *d_pipe_query = nullptr, *h_pipe_query = nullptr;        // [kPipeSlots][max_cols]: queries of tks_submit_host (device / pinned)
    uint32_t *h_pipe_res = nullptr;                                // [kPipeSlots][64 + 2 * kmax]: count | idx | val per slot, pinned
    uint32_t pipe_slot_k[kPipeSlots] = {};                         // k of the query in the slot
    uint64_t pipe_slot_ticket[kPipeSlots] = {};                    // ticket of the host-result query in the slot (0 = none)
    uint32_t pipe_seq = 0;                                         // sequence number of the last submitted query
    int pipe_last_slot = -1;
    int pipe_main_grid[4] = {0, 0, 0, 0};                          // grids of the 512-thread main kernels used here

    BscsrState *bs = nullptr;

    tks_stats stats{};

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define TKS_CUDA(h, call)                                                                              \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return (h)->fail(TKS_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),       \
                             __FILE__, __LINE__);                                                      \
    } while (0)

#ifdef __CUDACC__
// Launch with the programmatic-stream-serialization attribute: the grid may start before the previous kernel of the
// stream has finished (csr_topk.cuh: pdl_trigger / pdl_wait).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline bool pdl_enabled() {
    static const bool on = !(std::getenv("TKS_PDL") && std::atoi(std::getenv("TKS_PDL")) == 0);
    return on;
}

#endif

// api.cu: exclusive scan of n u32 values into n+1 u64 values on the handle's stream (setup only; synchronises)
int device_scan_u32(Handle *h, const uint32_t *d_in, uint64_t n, uint64_t *d_out);

// bscsr_api.cu
int bscsr_upload(Handle *h, uint32_t cols, uint32_t partitions, const uint64_t *packets_per_part,
                 const void *const *packets, const uint32_t *first_row, const uint64_t *nnz_per_part);
// GPU-side packer (bscsr_pack.cuh): row-sorted COO with raw ap_ufixed<32,1> values, host or device arrays
int bscsr_upload_coo(Handle *h, const uint32_t *row, const uint32_t *col, const uint32_t *val32, uint64_t nnz,
                     uint32_t num_rows, uint32_t cols, bool arrays_on_device);
int bscsr_set_query(Handle *h, const uint32_t *vec32_host, const uint32_t *vec32_dev, cudaStream_t s);
int bscsr_launch(Handle *h, cudaStream_t s);
int bscsr_submit_host(Handle *h, const uint32_t *vec32, uint32_t k, uint64_t *ticket);   // pipelined: sample of query i+1 beside query i
int bscsr_submit(Handle *h, const uint32_t *vec32_host, const uint32_t *vec32_dev, uint32_t k, cudaStream_t s,
                 bool query_ready, uint64_t *ticket);
int bscsr_fetch_ticket(Handle *h, uint64_t ticket, uint32_t *idx_out, uint32_t *val_out, uint32_t *count);
int bscsr_fetch(Handle *h);   // D2H of partition result words + host merge
int bscsr_read_result(Handle *h, uint32_t *idx_out, uint32_t *val_out, uint32_t k, uint32_t *count);
int bscsr_read_partition_results(Handle *h, uint32_t *idx_words, uint32_t *val_words);
int bscsr_partition_words_device(Handle *h, const uint32_t **d_words, uint32_t *n_words);
int bscsr_state_digest(Handle *h, uint64_t *digest, uint32_t n);
void bscsr_destroy(Handle *h);

}  // namespace tks

struct tks_handle : tks::Handle {};
