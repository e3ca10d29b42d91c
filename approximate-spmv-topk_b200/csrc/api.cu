// api.cu -- C ABI of libtopkspmv.so (see include/topkspmv.h), float CSR path and dispatch.
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "csr_build.cuh"
#include "csr_batched.cuh"
#include "csr_topk.cuh"
#include "handle.hpp"

using namespace tks;

static thread_local std::string g_create_error;

namespace {

constexpr uint32_t kKMax = 1024;
const int kCaps[4] = {256, 512, 1024, 2048};
// CTA size per CAP variant; the k <= 128 variant runs 18 warps per CTA (TKS_MAIN_THREADS=512 restores 16)
int kCapThreads[4] = {(std::getenv("TKS_MAIN_THREADS") && std::atoi(std::getenv("TKS_MAIN_THREADS")) == 512) ? 512 : 576, 512, 512, 256};

int cap_variant_for_k(uint32_t k) {
    if (k <= 128) return 0;
    if (k <= 384) return 1;
    if (k <= 896) return 2;
    return 3;
}

// CTA size of the main kernel inside a pipelined submit: 2 x 16 warps per SM at 56 registers leave 8192 registers per
// SM, exactly one 128-thread CTA of the sample kernel, which runs BESIDE the previous query's main kernel there
uint32_t env_u32(const char *name, uint32_t dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? (uint32_t)std::strtoul(v, nullptr, 10) : dflt;
}
// CTA size of the main kernel.  16-bit value modes work on 16 non-zeros per lane (~80 registers): 2 x 12 warps per SM.
int cap_threads(int variant, int vt) {
    static const int half_t = (int)env_u32("TKS_MAIN_THREADS_16BIT", 384u) / 32 * 32;   // A/B switch
    const int t = kCapThreads[variant];
    if (vt == 0) return t;
    const int ht = half_t < 128 ? 128 : (half_t > 384 ? 384 : half_t);
    return t < ht ? t : ht;
}
// ... inside a pipelined submit: leave room for one CTA of the sample kernel per SM (registers)
int pipe_threads(int variant, int vt) {
    static const int t0 = (int)env_u32("TKS_PIPE_THREADS", 512u) / 32 * 32;   // A/B switch: 256..576
    // 72 registers: 2 x 10 warps leave room for the sample CTAs in every SM sub-partition (with 2 x 11 they were measured
    // NOT to fit, r02i/j: the register file is allocated per sub-partition, 6 x 2304 of its 16384 leave one warp)
    static const int t16 = (int)env_u32("TKS_PIPE_THREADS_16BIT", 320u) / 32 * 32;
    if (vt != 0) { const int c = cap_threads(variant, vt); return c < t16 ? c : t16; }
    if (variant == 0) return t0 < 256 ? 256 : (t0 > 576 ? 576 : t0);
    return variant == 1 ? 448 : kCapThreads[variant];   // CAP 512 uses 64 registers: 2 x 14 warps leave 8192 for the sample CTA
}
int pipe_sample_threads(int vt) {
    static const int t = (int)env_u32("TKS_PIPE_SAMPLE_THREADS", 0u) / 32 * 32;
    if (t >= 32) return t > 256 ? 256 : t;
    (void)vt;
    return 128;   // the sample kernel is capped at 64 registers: 8192 per CTA; 64-thread CTAs made the 16-bit sample 200 us long (r02r)
}
// L1 / shared-memory split requested for every kernel of the float path, in percent of the maximum (-1: the driver's
// choice per kernel).  Kernels that share an SM in the pipelined path must agree on it: an SM cannot change the split
// while CTAs are resident.
int carveout_pct() { static const int v = std::getenv("TKS_CARVEOUT_PCT") ? std::atoi(std::getenv("TKS_CARVEOUT_PCT")) : -1; return v; }

size_t main_smem_bytes(uint32_t cols, int variant, int threads) {
    return (((cols + 1u) * 4u + 15u) & ~15u) + (size_t)(threads / 32) * kCaps[variant] * 8u;
}

// TKS_TMA=1: the k <= 128 main kernel streams through per-warp rings of bulk copies (cp.async.bulk + mbarrier) instead
// of register loads; chunk loads then start on 128-non-zero boundaries (in the sample kernel too)
// TKS_COL12=0: stream 16-bit column offsets even when 12 bits would do (A/B switch)
bool col12_enabled() { static const bool on = !(std::getenv("TKS_COL12") && std::atoi(std::getenv("TKS_COL12")) == 0); return on; }
uint32_t l2_prefetch_depth() { static const uint32_t v = env_u32("TKS_L2PF", 0u); return v > 8u ? 8u : v; }
bool tma_enabled() { static const bool on = std::getenv("TKS_TMA") && std::atoi(std::getenv("TKS_TMA")) != 0; return on; }
int tma_threads(int vt) { static const int t = (int)env_u32("TKS_TMA_THREADS", 0u) / 32 * 32; return t >= 64 ? t : (vt != 0 ? 384 : 448); }

// TKS_XCOPIES=1 (measurement variant): the k <= 128 main kernel keeps 32 interleaved copies of the query in shared memory
// (conflict-free gathers) and runs as ONE CTA per SM
bool xcopies_enabled() { static const bool on = std::getenv("TKS_XCOPIES") && std::atoi(std::getenv("TKS_XCOPIES")) != 0; return on; }
int xcopies_threads(int vt) {
    static const int t = (int)env_u32("TKS_XCOPIES_THREADS", 0u) / 32 * 32;
    const int cap = vt != 0 ? (int)kMainThreadsX16 : (int)kMainThreadsX;
    return (t >= 64 && t <= cap) ? t : cap;
}
size_t xcopies_smem_bytes(uint32_t cols, int threads) { return ((size_t)cols + 1u) * 128u + (size_t)(threads / 32) * 256u * 8u; }

// bounds of the device-side waits (peer records, pipelined hand-overs); raise them under compute-sanitizer
uint32_t spin_timeout_ms() { static const uint32_t v = env_u32("TKS_SPIN_TIMEOUT_MS", 2000u); return v; }
uint32_t tau_wait_us() { static const uint32_t v = env_u32("TKS_TAU_WAIT_US", 20000u); return v; }

inline bool half_mode(const Handle *h) { return h->cfg.value_type != TKS_VALUE_FP32; }   // 16-bit storage (half or bfloat16)
inline int value_type(const Handle *h) { return h->cfg.value_type; }

template <int CAP, int VT>
cudaError_t prep_main_t(Handle *h, int variant) {
    const int threads = cap_threads(variant, VT), pthreads = pipe_threads(variant, VT);
    size_t smem = main_smem_bytes(h->cfg.max_cols, variant, threads);
    cudaError_t e = cudaFuncSetAttribute(csr_topk_main_kernel<CAP, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    // Kernels of different queries share an SM in the pipelined path (sample and select CTAs beside main-kernel CTAs).
    // An SM cannot change its L1 / shared-memory split while CTAs are resident, so every kernel of the path asks for
    // the same split -- all shared memory (the matrix stream bypasses L1 anyway) -- or a CTA that needs a larger
    // carve-out than the resident kernel's would wait for the SM to drain.
    {
        // The split the main kernel asks for must leave room for what shares its SMs in a pipelined submit: two of its
        // own CTAs, two sample CTAs and the lean select CTA (left to itself the driver sizes the carve-out for the main
        // kernel's CTAs alone, and a sample CTA that does not fit waits for the SM to drain).  The rest stays L1, which
        // the stream needs for its loads in flight (with all of it turned into shared memory the kernel is 20 % slower).
        int pct = carveout_pct();
        if (pct < 0 && variant <= 1) {
            size_t ss = ((size_t)h->cfg.max_cols + 1u) * 4u;
            if (ss < (size_t)kHistScratchWords * 4u) ss = (size_t)kHistScratchWords * 4u;
            const size_t need = 2u * (main_smem_bytes(h->cfg.max_cols, variant, pthreads) + 1024u) + 2u * (ss + 1024u + 64u) +
                                (kSelectLeanDynSmem + sizeof(uint64_t) * kSelectSortCap + 2048u);
            const size_t total = 228u * 1024u;
            pct = (int)((need * 100u + total - 1u) / total);
            if (pct > 100) pct = 100;
        }
        if (pct >= 0) {
            e = cudaFuncSetAttribute(csr_topk_main_kernel<CAP, VT>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            if (e != cudaSuccess) return e;
            if constexpr (CAP == 256) {   // the instantiation that streams 12-bit column offsets
                e = cudaFuncSetAttribute(csr_topk_main_kernel<256, VT, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
                if (e != cudaSuccess) return e;
            }
        }
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_topk_main_kernel<CAP, VT>, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    h->main_grid[variant] = per_sm * h->num_sms;
    if constexpr (CAP == 256) {
        // the instantiation that streams 12-bit column offsets: same geometry, same attributes
        e = cudaFuncSetAttribute(csr_topk_main_kernel<256, VT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_topk_main_kernel<CAP, VT>, pthreads,
                                                      main_smem_bytes(h->cfg.max_cols, variant, pthreads));
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    h->pipe_main_grid[variant] = per_sm * h->num_sms;
    return cudaSuccess;
}

template <int VT>
cudaError_t prep_main_tma(Handle *h) {
    const int threads = tma_threads(VT);
    const size_t smem = main_smem_bytes(h->cfg.max_cols, 0, threads) + main_tma_extra_smem<VT>(threads / 32);
    cudaError_t e = cudaFuncSetAttribute(csr_topk_main_kernel<256, VT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_topk_main_kernel<256, VT, true>, threads, smem);
    if (e != cudaSuccess) return e;
    h->tma_grid = (per_sm < 1 ? 1 : per_sm) * h->num_sms;
    return cudaSuccess;
}

// Measured slower than the default kernel (r02ak: cfg2 main kernel 0.1939 -> 0.2083 ms, cfg2h 0.1479 -> 0.1563 ms): the
// 128 KB of copies push the carve-out to 196 KB, and the stream's loads in flight miss the L1 that is left (DESIGN.md
// section 4.1).  Compiled only with -DTKS_EXPERIMENT_XCOPIES (three more instantiations of the largest kernel).
#ifdef TKS_EXPERIMENT_XCOPIES
template <int VT>
cudaError_t prep_main_x(Handle *h) {
    const int threads = xcopies_threads(VT);
    const size_t smem = xcopies_smem_bytes(h->cfg.max_cols, threads);
    h->x_grid = 0;
    if (smem > 227u * 1024u) return cudaSuccess;   // too many columns for 32 copies: the default kernel runs
    cudaError_t e = cudaFuncSetAttribute(csr_topk_main_kernel<256, VT, false, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    h->x_grid = h->num_sms;
    return cudaSuccess;
}
#endif

template <int CAP>
cudaError_t prep_main(Handle *h, int variant) {
#ifdef TKS_EXPERIMENT_XCOPIES
    if (CAP == 256 && xcopies_enabled()) {
        cudaError_t e = value_type(h) == TKS_VALUE_FP16 ? prep_main_x<1>(h) : value_type(h) == TKS_VALUE_BF16 ? prep_main_x<2>(h) : prep_main_x<0>(h);
        if (e != cudaSuccess) return e;
    }
#endif
    if (CAP == 256 && tma_enabled()) {
        cudaError_t e = value_type(h) == TKS_VALUE_FP16 ? prep_main_tma<1>(h) : value_type(h) == TKS_VALUE_BF16 ? prep_main_tma<2>(h) : prep_main_tma<0>(h);
        if (e != cudaSuccess) return e;
    }
    switch (value_type(h)) {
        case TKS_VALUE_FP16: return prep_main_t<CAP, 1>(h, variant);
        case TKS_VALUE_BF16: return prep_main_t<CAP, 2>(h, variant);
        default: return prep_main_t<CAP, 0>(h, variant);
    }
}

// seq != 0: a pipelined submit (512-thread CTAs for the k <= 128 variant, hand-over by sequence numbers)
template <int CAP>
void launch_main(Handle *h, int variant, const CsrDevice &m, const float *x, RunState *st, uint64_t *pool, uint32_t k,
                 cudaStream_t s, bool pdl, uint32_t seq = 0, uint64_t *stamp = nullptr) {
    const int threads = seq ? pipe_threads(variant, value_type(h)) : cap_threads(variant, value_type(h));
    size_t smem = main_smem_bytes(m.cols, variant, threads);
    const int tie_higher = h->cfg.tie_break == TKS_TIE_HIGHER_INDEX;
    const dim3 grid(seq ? h->pipe_main_grid[variant] : h->main_grid[variant]), block(threads);
    const uint32_t tw = tau_wait_us();
    if (CAP == 256 && tma_enabled()) {
        const int tt = tma_threads(value_type(h));
        const dim3 tgrid(h->tma_grid), tblock(tt);
        const size_t base = main_smem_bytes(m.cols, 0, tt);
        switch (value_type(h)) {
            case TKS_VALUE_FP16: launch_pdl(csr_topk_main_kernel<256, 1, true>, tgrid, tblock, base + main_tma_extra_smem<1>(tt / 32), s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
            case TKS_VALUE_BF16: launch_pdl(csr_topk_main_kernel<256, 2, true>, tgrid, tblock, base + main_tma_extra_smem<2>(tt / 32), s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
            default: launch_pdl(csr_topk_main_kernel<256, 0, true>, tgrid, tblock, base + main_tma_extra_smem<0>(tt / 32), s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
        }
        return;
    }
#ifdef TKS_EXPERIMENT_XCOPIES
    if (CAP == 256 && h->d_col12 && xcopies_enabled() && h->x_grid) {
        const int xt = xcopies_threads(value_type(h));
        const dim3 xgrid(h->x_grid), xblock(xt);
        const size_t xs = xcopies_smem_bytes(m.cols, xt);
        switch (value_type(h)) {
            case TKS_VALUE_FP16: launch_pdl(csr_topk_main_kernel<256, 1, false, true, 5>, xgrid, xblock, xs, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
            case TKS_VALUE_BF16: launch_pdl(csr_topk_main_kernel<256, 2, false, true, 5>, xgrid, xblock, xs, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
            default: launch_pdl(csr_topk_main_kernel<256, 0, false, true, 5>, xgrid, xblock, xs, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
        }
        return;
    }
#endif
    if (CAP == 256 && h->d_col12) {
        switch (value_type(h)) {
            case TKS_VALUE_FP16: launch_pdl(csr_topk_main_kernel<256, 1, false, true>, grid, block, smem, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
            case TKS_VALUE_BF16: launch_pdl(csr_topk_main_kernel<256, 2, false, true>, grid, block, smem, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
            default: launch_pdl(csr_topk_main_kernel<256, 0, false, true>, grid, block, smem, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
        }
        return;
    }
    switch (value_type(h)) {
        case TKS_VALUE_FP16: launch_pdl(csr_topk_main_kernel<CAP, 1>, grid, block, smem, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
        case TKS_VALUE_BF16: launch_pdl(csr_topk_main_kernel<CAP, 2>, grid, block, smem, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
        default: launch_pdl(csr_topk_main_kernel<CAP, 0>, grid, block, smem, s, pdl, m, x, st, pool, k, tie_higher, seq, tw, stamp); break;
    }
}

void launch_main_variant(Handle *h, int variant, const CsrDevice &m, const float *x, RunState *st, uint64_t *pool,
                         uint32_t k, cudaStream_t s, bool pdl, uint32_t seq = 0, uint64_t *stamp = nullptr) {
    switch (variant) {
        case 0: launch_main<256>(h, 0, m, x, st, pool, k, s, pdl, seq, stamp); break;
        case 1: launch_main<512>(h, 1, m, x, st, pool, k, s, pdl, seq, stamp); break;
        case 2: launch_main<1024>(h, 2, m, x, st, pool, k, s, pdl, seq, stamp); break;
        default: launch_main<2048>(h, 3, m, x, st, pool, k, s, pdl, seq, stamp); break;
    }
}

CsrDevice csr_device(const Handle *h) {
    return CsrDevice{h->d_val, h->d_col16, h->d_col12, reinterpret_cast<const uint8_t *>(h->d_rowbits), h->d_chunk_start,
                     h->d_chunk_ord, h->d_row_map, h->n_chunks, h->cols, (uint32_t)h->row_offset,
                     (uint32_t)value_type(h), tma_enabled() ? 128u : (value_type(h) != 0 ? 16u : 8u), l2_prefetch_depth()};
}

__global__ void widen_u32_to_u64_kernel(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

uint64_t max_pool_keys(const Handle *h) {
    uint64_t best = 0;
    for (int v = 0; v < 4; v++) {
        uint64_t n = (uint64_t)h->main_grid[v] * (cap_threads(v, value_type(h)) / 32) * kCaps[v];
        const uint64_t np = (uint64_t)h->pipe_main_grid[v] * (pipe_threads(v, value_type(h)) / 32) * kCaps[v];
        if (np > n) n = np;
        if (n > best) best = n;
    }
    return best;
}

int alloc_query_side(Handle *h) {
    const uint32_t mb = (uint32_t)h->cfg.max_batch;
    h->kmax = kKMax;
    TKS_CUDA(h, cudaMalloc(&h->d_x, (size_t)mb * h->cfg.max_cols * sizeof(float)));
    TKS_CUDA(h, cudaMalloc(&h->d_state, mb * sizeof(RunState)));
    TKS_CUDA(h, cudaMemset(h->d_state, 0, mb * sizeof(RunState)));
    h->pool_cap = max_pool_keys(h);
    TKS_CUDA(h, cudaMalloc(&h->d_pool, h->pool_cap * sizeof(uint64_t)));
    h->n_sample_cap = 8192;
    TKS_CUDA(h, cudaMalloc(&h->d_sample_keys, h->n_sample_cap * sizeof(uint32_t)));
    TKS_CUDA(h, cudaMalloc(&h->d_res_keys, (size_t)mb * h->kmax * sizeof(uint64_t)));
    // results live in ONE block [count: mb (padded)] [idx: mb x kmax] [val: mb x kmax], mirrored in pinned host
    // memory, so that a run's results come back with a single device-to-host copy
    const size_t cnt_words = ((size_t)mb + 63u) & ~(size_t)63u;
    h->res_block_bytes = (cnt_words + 2 * (size_t)mb * h->kmax) * 4u;
    TKS_CUDA(h, cudaMalloc(&h->d_res_block, h->res_block_bytes));
    TKS_CUDA(h, cudaMemset(h->d_res_block, 0, h->res_block_bytes));
    TKS_CUDA(h, cudaMallocHost(&h->h_res_block, h->res_block_bytes));
    h->d_res_count = h->d_res_block;
    h->d_res_idx = h->d_res_block + cnt_words;
    h->d_res_val = reinterpret_cast<float *>(h->d_res_idx + (size_t)mb * h->kmax);
    h->h_res_count = h->h_res_block;
    h->h_res_idx = h->h_res_block + cnt_words;
    h->h_res_val = reinterpret_cast<float *>(h->h_res_idx + (size_t)mb * h->kmax);
    TKS_CUDA(h, cudaMallocHost(&h->h_x, (size_t)mb * h->cfg.max_cols * sizeof(float)));
    if (mb > 1 && h->batched_ok) {
        const uint32_t npass = (mb + kBqPerPass - 1) / kBqPerPass;
        h->bpool_cap = h->cfg.batch_pool_cap > 0 ? (uint32_t)h->cfg.batch_pool_cap : 32768u;
        h->b_sample_cap = 8192;
        TKS_CUDA(h, cudaMalloc(&h->d_xT, (size_t)npass * batched_table_bytes((uint32_t)h->cfg.max_cols)));
        TKS_CUDA(h, cudaMalloc(&h->d_bpool, (size_t)mb * h->bpool_cap * sizeof(uint64_t)));
        TKS_CUDA(h, cudaMalloc(&h->d_pass_counter, npass * sizeof(uint32_t)));
        TKS_CUDA(h, cudaMemset(h->d_pass_counter, 0, npass * sizeof(uint32_t)));
        TKS_CUDA(h, cudaMalloc(&h->d_bsample_keys, (size_t)mb * h->b_sample_cap * sizeof(uint32_t)));
    }
    return TKS_OK;
}

void drop_run_graph(Handle *h) {
    if (h->run_graph) cudaGraphExecDestroy(h->run_graph);
    h->run_graph = nullptr;
    h->run_graph_k = 0;
}

int pipe_drain(Handle *h);

void free_matrix(Handle *h) {
    // a matrix is replaced: nothing submitted against the old one may still be running
    if (h->d_pipe_state) { pipe_drain(h); cudaDeviceSynchronize(); }
    drop_run_graph(h);
    cudaFree(h->d_val); h->d_val = nullptr;
    cudaFree(h->d_col16); h->d_col16 = nullptr;
    cudaFree(h->d_col12); h->d_col12 = nullptr;
    cudaFree(h->d_rowbits); h->d_rowbits = nullptr;
    cudaFree(h->d_ptr64); h->d_ptr64 = nullptr;
    cudaFree(h->d_chunk_start); h->d_chunk_start = nullptr;
    cudaFree(h->d_chunk_ord); h->d_chunk_ord = nullptr;
    cudaFree(h->d_row_map); h->d_row_map = nullptr;
    h->have_matrix = false;
}

// Build colf + chunk table (+ row map when rows are empty) from a device CSR.  val is copied unless adopted.
template <typename P>
int build_from_device_csr(Handle *h, uint64_t rows, uint32_t cols, uint64_t nnz, const P *d_ptr,
                          const uint32_t *d_idx, const float *d_val_src, float *d_val_adopt) {
    cudaStream_t s = h->stream;
    const size_t pad = 2048;   // over-read slack of the 256-bit streaming loads (zero filled)
    // work units: cfg.chunk_nnz if given; else sized below for large streams, 4096 non-zeros for small matrices.  Every
    // unit start costs a warp three dependent round trips (scheduler atomic, chunk table, first loads), ~2 us: cfg2 steps
    // with 2048 / 4096 / 8192 non-zeros per unit take 0.283 / 0.201 / 0.191 ms (r02af).  A tail of quarter-size units
    // (TKS_CHUNK_TAIL=1), meant to let the persistent warps run dry together, costs more than it balances: 0.2041 vs
    // 0.1915 ms (r02ag).
    uint32_t n_big = 0, chunk_small = 0;
    if (h->cfg.chunk_nnz > 0) {
        h->chunk_nnz = (uint32_t)h->cfg.chunk_nnz;
    } else {
        uint32_t dflt = 4096u;
        if (h->cfg.max_batch == 1 && nnz >= (32ull << 20)) {
            // The persistent warps run at the same speed, so the dynamic scheduler ends up handing every warp
            // ceil(units / warps) units: the main kernel alone takes 0.1966 / 0.1925 / 0.2022 / 0.1923 / 0.2134 / 0.1968 ms
            // with 8192 / 9216 / 11264 / 12288 / 16384 / 20480 non-zeros per unit on cfg2 -- exactly the order of
            // ceil(units / 5328 warps) x unit (r02aq).  Hence: a whole number m of units per resident warp of the k <= 128
            // main kernel, the largest m that keeps a unit at ~48 warp iterations or more (12 000 non-zeros; 24 000
            // with 16-bit values, whose iterations are 512 non-zeros).
            const uint64_t warps = (uint64_t)(h->main_grid[0] > 0 ? h->main_grid[0] : 2 * h->num_sms) *
                                   (uint64_t)(cap_threads(0, value_type(h)) / 32);
            const uint64_t per_warp = (nnz + warps - 1) / warps;
            const uint64_t tmin = half_mode(h) ? 24000u : 12000u;
            const uint64_t m = per_warp / tmin > 0 ? per_warp / tmin : 1u;
            uint64_t unit = (per_warp + m - 1) / m;
            if (unit < 4096u) unit = 4096u;
            dflt = (uint32_t)unit;   // rounded up to whole warp iterations below
        } else if (h->cfg.max_batch > 1 && nnz >= (32ull << 20)) {
            // batched handles: the same rule over the batched kernel's streams (one CTA per SM, eight quads per warp, a
            // warp takes eight units at a time and waits for the longest), at ~4096 non-zeros per unit -- eight streams
            // per warp want smaller units to balance (cfg5 main kernel 10.85 ms with 4096, 11.42 ms with 8192, r02ah; with
            // the whole-number rule 10.56 ms, and 1.73 -> 1.39 ms on one rank's share of 8 GPUs, r02as)
            const uint64_t streams = (uint64_t)h->num_sms * (kBThreads / 32u) * kBStreams;
            const uint64_t per_stream = (nnz + streams - 1) / streams;
            const uint64_t m = per_stream / 4096u > 0 ? per_stream / 4096u : 1u;
            const uint64_t unit = (per_stream + m - 1) / m;
            dflt = (uint32_t)(unit < 2048u ? 2048u : unit);
        }
        h->chunk_nnz = env_u32("TKS_CHUNK_NNZ", dflt);   // env: A/B switch
        if (nnz >= (32ull << 20) && env_u32("TKS_CHUNK_TAIL", 0u)) chunk_small = h->chunk_nnz / 4u;
    }
    h->chunk_nnz = (h->chunk_nnz + kElemsPerIter - 1) / kElemsPerIter * kElemsPerIter;
    uint64_t nch = (nnz + h->chunk_nnz - 1) / h->chunk_nnz;
    if (chunk_small) {
        chunk_small = (chunk_small + kElemsPerIter - 1) / kElemsPerIter * kElemsPerIter;
        n_big = (uint32_t)(nnz / 10u * 9u / h->chunk_nnz);
        const uint64_t rest = nnz - (uint64_t)n_big * h->chunk_nnz;
        nch = n_big + (rest + chunk_small - 1) / chunk_small;
    } else {
        n_big = (uint32_t)(nch > 0xFFFFFFF0ull ? 0u : nch);
        chunk_small = h->chunk_nnz;
    }
    if (nch == 0) nch = 1;
    if (nch > 0xFFFFFFF0ull) return h->fail(TKS_EINVAL, "too many chunks (%llu)", (unsigned long long)nch);
    h->n_chunks = (uint32_t)nch;
    h->stats.work_unit_nnz = h->chunk_nnz;
    h->stats.work_units = h->n_chunks;

    if (half_mode(h)) {
        // the reference's half-precision mode (host_spmv_topk_csr_gpu.cu:133,152: float_to_half of every value)
        const float *src = d_val_adopt ? d_val_adopt : d_val_src;
        TKS_CUDA(h, cudaMalloc(&h->d_val, nnz * sizeof(__half) + pad));
        TKS_CUDA(h, cudaMemsetAsync(reinterpret_cast<uint8_t *>(h->d_val) + nnz * sizeof(__half), 0, pad, s));
        if (nnz > 0 && value_type(h) == TKS_VALUE_FP16)
            csr_vals_to_half_kernel<<<h->num_sms * 8, 256, 0, s>>>(src, nnz, reinterpret_cast<__half *>(h->d_val));
        if (nnz > 0 && value_type(h) == TKS_VALUE_BF16)
            csr_vals_to_bf16_kernel<<<h->num_sms * 8, 256, 0, s>>>(src, nnz, reinterpret_cast<__nv_bfloat16 *>(h->d_val));
        if (d_val_adopt) {
            TKS_CUDA(h, cudaStreamSynchronize(s));
            cudaFree(d_val_adopt);
        }
    } else if (d_val_adopt) {
        h->d_val = d_val_adopt;
    } else {
        TKS_CUDA(h, cudaMalloc(&h->d_val, nnz * sizeof(float) + pad));
        TKS_CUDA(h, cudaMemsetAsync(reinterpret_cast<uint8_t *>(h->d_val) + nnz * sizeof(float), 0, pad, s));
        TKS_CUDA(h, cudaMemcpyAsync(h->d_val, d_val_src, nnz * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    TKS_CUDA(h, cudaMalloc(&h->d_col16, nnz * sizeof(uint16_t) + pad));
    TKS_CUDA(h, cudaMemsetAsync(reinterpret_cast<uint8_t *>(h->d_col16) + nnz * sizeof(uint16_t), 0, pad, s));
    const size_t rowbits_bytes = ((nnz + 31) / 32) * 4 + pad;
    TKS_CUDA(h, cudaMalloc(&h->d_rowbits, rowbits_bytes));
    TKS_CUDA(h, cudaMemsetAsync(h->d_rowbits, 0, rowbits_bytes, s));
    TKS_CUDA(h, cudaMalloc(&h->d_chunk_start, (nch + 1) * sizeof(uint64_t)));
    TKS_CUDA(h, cudaMalloc(&h->d_chunk_ord, nch * sizeof(uint32_t)));
    uint32_t *d_err = nullptr, *d_flag = nullptr;
    uint64_t *d_ord = nullptr;
    TKS_CUDA(h, cudaMalloc(&d_err, sizeof(uint32_t)));
    TKS_CUDA(h, cudaMemsetAsync(d_err, 0, sizeof(uint32_t), s));
    TKS_CUDA(h, cudaMalloc(&d_flag, (rows ? rows : 1) * sizeof(uint32_t)));
    TKS_CUDA(h, cudaMalloc(&d_ord, (rows + 1) * sizeof(uint64_t)));

    if (nnz > 0) csr_copy_cols_kernel<P><<<h->num_sms * 8, 256, 0, s>>>(d_idx, nnz, cols, h->d_col16, d_err);
    if (cols <= 1024 && col12_enabled() && !tma_enabled()) {
        // columns <= 1023: column * 4 fits 12 bits -- the stream the k <= 128 single-query kernels read (1.5 B / non-zero)
        const size_t c12_bytes = (nnz + 7) / 8 * 12 + pad;
        TKS_CUDA(h, cudaMalloc(&h->d_col12, c12_bytes));
        TKS_CUDA(h, cudaMemsetAsync(h->d_col12, 0, c12_bytes, s));
        if (nnz > 0) csr_pack_cols12_kernel<<<h->num_sms * 8, 256, 0, s>>>(h->d_col16, nnz, h->d_col12);
    }
    if (rows > 0) {
        csr_mark_rows_kernel<P><<<(uint32_t)((rows + 255) / 256), 256, 0, s>>>(d_ptr, rows, nnz, h->d_rowbits, d_flag, d_err);
        int rc = device_scan_u32(h, d_flag, rows, d_ord);
        if (rc) return rc;
    } else {
        TKS_CUDA(h, cudaMemsetAsync(d_ord, 0, sizeof(uint64_t), s));
    }
    uint64_t n_nonempty = 0;
    TKS_CUDA(h, cudaMemcpyAsync(&n_nonempty, d_ord + rows, 8, cudaMemcpyDeviceToHost, s));
    TKS_CUDA(h, cudaStreamSynchronize(s));
    const bool has_empty = n_nonempty != rows;
    if (has_empty) {
        TKS_CUDA(h, cudaMalloc(&h->d_row_map, (n_nonempty ? n_nonempty : 1) * sizeof(uint32_t)));
        csr_row_map_kernel<<<(uint32_t)((rows + 255) / 256), 256, 0, s>>>(d_flag, d_ord, rows, h->d_row_map);
    }
    csr_chunk_table_kernel<P><<<(uint32_t)((nch + 1 + 127) / 128), 128, 0, s>>>(
        d_ptr, rows, nnz, h->chunk_nnz, n_big, chunk_small, h->n_chunks, has_empty ? d_ord : nullptr, h->d_chunk_start,
        h->d_chunk_ord);
    uint32_t herr = 0;
    TKS_CUDA(h, cudaMemcpyAsync(&herr, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    TKS_CUDA(h, cudaStreamSynchronize(s));
    TKS_CUDA(h, cudaGetLastError());
    cudaFree(d_err); cudaFree(d_flag); cudaFree(d_ord);
    if (herr) {
        free_matrix(h);
        return h->fail(TKS_EINVAL, "invalid CSR:%s%s", (herr & kErrColRange) ? " column index >= cols;" : "",
                       (herr & kErrPtrOrder) ? " row_ptr not monotone, out of range, or not running from 0 to nnz;" : "");
    }
    h->rows = rows; h->cols = cols; h->nnz = nnz;
    // bytes one k <= 128 query streams: values + column offsets (2 bytes, or 1.5 when packed to 12 bits) + row-start bits + tables
    h->device_bytes = nnz * (half_mode(h) ? 2ull : 4ull) + (h->d_col12 ? nnz * 3ull / 2ull : nnz * 2ull) + (nnz + 7) / 8 +
                      (nch + 1) * 8ull + nch * 4ull + (has_empty ? n_nonempty * 4ull : 0ull);
    h->have_matrix = true;
    h->have_result = false;
    h->stats.rows = rows; h->stats.cols = cols; h->stats.nnz = nnz; h->stats.packets = 0;
    h->stats.device_bytes = h->device_bytes;
    return TKS_OK;
}

int check_float_upload(Handle *h, uint64_t rows, uint32_t cols, uint64_t nnz, uint64_t row_offset) {
    if (h->cfg.mode != TKS_MODE_FLOAT_CSR) return h->fail(TKS_ESTATE, "handle is not in FLOAT_CSR mode");
    if (cols == 0 || cols > (uint32_t)h->cfg.max_cols)
        return h->fail(TKS_EINVAL, "cols=%u outside 1..max_cols=%d", cols, h->cfg.max_cols);
    if (rows + row_offset > 0xFFFFFFFFull) return h->fail(TKS_EINVAL, "row ids exceed 32 bits");
    (void)nnz;
    return TKS_OK;
}

// the k <= 128 kernels stream the 12-bit column offsets when the matrix has them
bool use_col12(const Handle *h, uint32_t k) { return h->d_col12 != nullptr && k <= 128; }

// The sample kernel of one query: one warp per resident warp slot of the main kernel's grid; large shards get more
// warps (up to the key buffer's capacity) so that the ~1 % sample stays a few iterations deep instead of a long
// latency-bound walk.  threads: CTA size (256 alone on the device, 128 beside a running main kernel).
void launch_sample(Handle *h, const CsrDevice &m, const float *x, RunState *st, uint32_t *sample_keys, uint32_t k,
                   cudaStream_t s, uint32_t threads, uint32_t seq, uint64_t *stamp = nullptr) {
    const uint32_t epi = elems_per_iter(value_type(h));   // non-zeros per warp iteration: 256 (fp32) or 512 (16-bit values)
    uint64_t want = (uint64_t)h->num_sms * 32u;
    const uint64_t for_depth4 = h->nnz / 100u / (4u * epi);
    if (for_depth4 > want) want = for_depth4;
    if (want > h->n_sample_cap) want = h->n_sample_cap;
    uint32_t n_sample = h->n_chunks < want ? h->n_chunks : (uint32_t)want;
    size_t sample_smem = ((size_t)h->cols + 1u) * 4u;
    if (sample_smem < (size_t)kHistScratchWords * 4u) sample_smem = (size_t)kHistScratchWords * 4u;
    const uint32_t sgrid = (n_sample * kWarp + threads - 1) / threads;
    // the sample grows with the matrix (~1 % of the non-zeros) so that the candidates stay a few thousand
    uint64_t si = (h->nnz / 100u + (uint64_t)n_sample * epi - 1) / ((uint64_t)n_sample * epi);
    const uint32_t max_si = h->chunk_nnz / epi;
    const uint32_t sample_iters = (uint32_t)(si < 2 ? 2 : (si > max_si ? (max_si < 2 ? 2 : max_si) : si));
    if (use_col12(h, k)) {
        switch (value_type(h)) {
            case TKS_VALUE_FP16: csr_sample_kernel<1, true><<<sgrid, threads, sample_smem, s>>>(m, x, st, sample_keys, n_sample, sample_iters, k, seq, stamp); break;
            case TKS_VALUE_BF16: csr_sample_kernel<2, true><<<sgrid, threads, sample_smem, s>>>(m, x, st, sample_keys, n_sample, sample_iters, k, seq, stamp); break;
            default: csr_sample_kernel<0, true><<<sgrid, threads, sample_smem, s>>>(m, x, st, sample_keys, n_sample, sample_iters, k, seq, stamp); break;
        }
        return;
    }
    switch (value_type(h)) {
        case TKS_VALUE_FP16: csr_sample_kernel<1><<<sgrid, threads, sample_smem, s>>>(m, x, st, sample_keys, n_sample, sample_iters, k, seq, stamp); break;
        case TKS_VALUE_BF16: csr_sample_kernel<2><<<sgrid, threads, sample_smem, s>>>(m, x, st, sample_keys, n_sample, sample_iters, k, seq, stamp); break;
        default: csr_sample_kernel<0><<<sgrid, threads, sample_smem, s>>>(m, x, st, sample_keys, n_sample, sample_iters, k, seq, stamp); break;
    }
}

// sample -> main -> select for query q alone (also the fallback of a batched query whose pool overflowed)
void launch_single_query(Handle *h, uint32_t q, uint32_t k, cudaStream_t s, bool profile, bool to_host = false,
                         const PeerExchange *px = nullptr, uint32_t seq = 0) {
    const int variant = cap_variant_for_k(k);
    const int tie_higher = h->cfg.tie_break == TKS_TIE_HIGHER_INDEX;
    const CsrDevice m = csr_device(h);
    const float *x = h->d_x + (size_t)q * h->cols;
    RunState *st = h->d_state + q;
    launch_sample(h, m, x, st, h->d_sample_keys, k, s, kSampleThreads, 0u);
    // the three kernels of a query overlap their launch and set-up with the previous one's tail (programmatic
    // dependent launch); not while the dominant kernel is being timed alone
    const bool pdl = pdl_enabled() && !profile;
    if (profile) cudaEventRecord(h->evm0, s);
    launch_main_variant(h, variant, m, x, st, h->d_pool, k, s, pdl);
    if (profile) cudaEventRecord(h->evm1, s);
    // to_host (blocking tks_run, one query): indices, scores and the count go straight into the pinned host block the
    // caller reads (zero-copy stores, 1.2 KB), which saves the device-to-host copy after the kernel; the keys stay in HBM
    uint32_t *o_idx = (to_host ? h->h_res_idx : h->d_res_idx) + (size_t)q * h->kmax;
    float *o_val = (to_host ? h->h_res_val : h->d_res_val) + (size_t)q * h->kmax;
    uint32_t *o_cnt = (to_host ? h->h_res_count : h->d_res_count) + q;
    h->res_on_host = to_host;
    if (px && px->world > 1) {
        // several GPUs: local select + exchange over the peer windows + merge in this one launch
        launch_pdl(select_topk_kernel<true>, dim3(1), dim3(kSelectThreads), (size_t)kSelectDynSmem, s, pdl,
                   (const uint64_t *)h->d_pool, 0u, st, 0u, 0u, k, tie_higher, h->d_res_keys + (size_t)q * h->kmax,
                   o_idx, o_val, 0u, o_cnt, (uint32_t *)nullptr, *px, seq, 0u, 0u, (uint64_t *)nullptr, kSelectSmemKeys);
        return;
    }
    launch_pdl(select_topk_kernel<false>, dim3(1), dim3(kSelectThreads), (size_t)kSelectDynSmem, s, pdl,
               (const uint64_t *)h->d_pool, 0u, st, 0u, 0u, k, tie_higher, h->d_res_keys + (size_t)q * h->kmax,
               o_idx, o_val, 0u, o_cnt, (uint32_t *)nullptr, PeerExchange{}, 0u, 0u, 0u, (uint64_t *)nullptr, kSelectSmemKeys);
}

// ---- pipelined submits -------------------------------------------------------------------------------------------

int pipe_init(Handle *h) {
    if (h->d_pipe_state) return TKS_OK;
    {
        const int want = (int)env_u32("TKS_PIPE_SLOTS", (uint32_t)Handle::kPipeSlots);
        h->pipe_slots = want < 2 ? 2 : (want > Handle::kPipeSlots ? Handle::kPipeSlots : want);
    }
    int lo = 0, hi = 0;
    TKS_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = greatest priority
    // The block scheduler serves pending grids by priority, then age, and does not look past one whose CTAs do not fit:
    // the select CTA of the previous query (too large to sit beside two main-kernel CTAs, so it is pending until one
    // retires) must not stand in front of the sample CTAs of the next query (small, they fit at once).  Hence sample >
    // select > the caller's stream (main kernels).
    int sel = hi + (int)env_u32("TKS_PIPE_SELECT_PRIO_DELTA", 1u);
    if (sel > lo) sel = lo;
    TKS_CUDA(h, cudaStreamCreateWithPriority(&h->pipe_sample_stream, cudaStreamNonBlocking, hi));
    TKS_CUDA(h, cudaStreamCreateWithPriority(&h->pipe_select_stream, cudaStreamNonBlocking, sel));
    for (int i = 0; i < Handle::kPipeSlots; i++) {
        TKS_CUDA(h, cudaEventCreateWithFlags(&h->pipe_ev_done[i], cudaEventDisableTiming));
        TKS_CUDA(h, cudaMalloc(&h->d_pipe_pool[i], h->pool_cap * sizeof(uint64_t)));
    }
    TKS_CUDA(h, cudaEventCreateWithFlags(&h->pipe_ev_query, cudaEventDisableTiming));
    TKS_CUDA(h, cudaMalloc(&h->d_pipe_sample_keys, h->n_sample_cap * sizeof(uint32_t)));
    const size_t stamp_bytes = (size_t)Handle::kPipeStamps * kStampWords * sizeof(uint64_t);
    TKS_CUDA(h, cudaMalloc(&h->d_pipe_stamps, stamp_bytes));
    TKS_CUDA(h, cudaMemset(h->d_pipe_stamps, 0, stamp_bytes));
    TKS_CUDA(h, cudaMallocHost(&h->h_pipe_stamps, stamp_bytes));
    const size_t qbytes = (size_t)Handle::kPipeSlots * h->cfg.max_cols * sizeof(float);
    TKS_CUDA(h, cudaMalloc(&h->d_pipe_query, qbytes));
    TKS_CUDA(h, cudaMallocHost(&h->h_pipe_query, qbytes));
    TKS_CUDA(h, cudaMallocHost(&h->h_pipe_res, (size_t)Handle::kPipeSlots * (64u + 2u * (size_t)h->kmax) * 4u));
    TKS_CUDA(h, cudaMalloc(&h->d_pipe_state, Handle::kPipeSlots * sizeof(RunState)));
    TKS_CUDA(h, cudaMemset(h->d_pipe_state, 0, Handle::kPipeSlots * sizeof(RunState)));
    return TKS_OK;
}

// Block the host until every pipelined query in flight has been selected (the un-pipelined entry points share the
// result block with them).
int pipe_drain(Handle *h) {
    for (int i = 0; i < Handle::kPipeSlots; i++) {
        if (!h->pipe_busy[i]) continue;
        TKS_CUDA(h, cudaEventSynchronize(h->pipe_ev_done[i]));
        h->pipe_busy[i] = false;
    }
    return TKS_OK;
}

template <bool SAMPLE>
void launch_batched_kernel(Handle *h, const CsrDevice &m, const BatchedArgs &a, uint32_t grid, cudaStream_t s) {
    const size_t smem = batched_smem_bytes(m.cols);
    if (h->cfg.batch_fma) csr_batched_kernel<SAMPLE, true><<<grid, kBThreads, smem, s>>>(m, a);
    else csr_batched_kernel<SAMPLE, false><<<grid, kBThreads, smem, s>>>(m, a);
}

// One matrix pass per 32 queries (csr_batched.cuh).
void launch_batched(Handle *h, uint32_t k, cudaStream_t s, bool profile) {
    const CsrDevice m = csr_device(h);
    BatchedArgs a{};
    a.xT = h->d_xT;
    a.st = h->d_state;
    a.pool = h->d_bpool;
    a.pool_cap = h->bpool_cap;
    a.batch = h->batch;
    a.npass = (h->batch + kBqPerPass - 1) / kBqPerPass;
    a.pass_counter = h->d_pass_counter;
    a.sample_keys = h->d_bsample_keys;
    a.n_sample = h->n_chunks < h->b_sample_cap ? h->n_chunks : h->b_sample_cap;
    a.stride = h->n_chunks / a.n_sample;
    {
        uint64_t sb = (h->nnz / 50u + (uint64_t)a.n_sample * kBStage - 1) / ((uint64_t)a.n_sample * kBStage);
        a.sample_batches = (uint32_t)(sb < 16 ? 16 : (sb > 128 ? 128 : sb));   // batches of kBStage = 32 non-zeros
    }
    a.tie_higher = h->cfg.tie_break == TKS_TIE_HIGHER_INDEX;
    batched_transpose_kernel<<<h->num_sms, 256, 0, s>>>(h->d_x, h->batch, h->cols, a.npass, h->d_xT, value_type(h));
    const uint32_t warps_per_cta = kBThreads / kWarp;
    uint32_t sgrid = ((a.n_sample + kBStreams - 1u) / kBStreams + warps_per_cta - 1) / warps_per_cta;
    if (sgrid > (uint32_t)h->num_sms) sgrid = (uint32_t)h->num_sms;
    launch_batched_kernel<true>(h, m, a, sgrid, s);
    batched_tau_kernel<<<h->batch, 256, (size_t)a.n_sample * 4u, s>>>(a, k);
    if (profile) cudaEventRecord(h->evm0, s);
    launch_batched_kernel<false>(h, m, a, (uint32_t)h->num_sms, s);
    if (profile) cudaEventRecord(h->evm1, s);
    select_topk_kernel<false><<<h->batch, kSelectThreads, kSelectDynSmem, s>>>(
        h->d_bpool, h->bpool_cap, h->d_state, 0u, h->bpool_cap, k, a.tie_higher, h->d_res_keys, h->d_res_idx,
        h->d_res_val, h->kmax, h->d_res_count, h->d_pass_counter, PeerExchange{}, 0u, 0u, 0u, nullptr, kSelectSmemKeys);
    h->res_on_host = false;
}

bool graphs_enabled() {
    static const bool on = !(std::getenv("TKS_GRAPH") && std::atoi(std::getenv("TKS_GRAPH")) == 0);
    return on;
}

// The three launches of one blocking query (results straight to the pinned host block) as ONE graph launch: captured
// from the very calls launch_single_query makes, programmatic-dependent-launch edges included, once per k.  Returns
// false when the capture is not possible (the caller then launches the kernels one by one).
bool launch_query_graph(Handle *h, uint32_t k, cudaStream_t s) {
    if (h->run_graph_failed || !graphs_enabled()) return false;
    if (!h->run_graph || h->run_graph_k != k) {
        drop_run_graph(h);
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); h->run_graph_failed = true; return false; }
        launch_single_query(h, 0, k, s, false, true);
        const cudaError_t e1 = cudaStreamEndCapture(s, &g);
        cudaError_t e2 = cudaErrorUnknown;
        if (e1 == cudaSuccess && g) e2 = cudaGraphInstantiate(&h->run_graph, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            cudaGetLastError();
            h->run_graph = nullptr;
            h->run_graph_failed = true;
            return false;
        }
        h->run_graph_k = k;
    }
    if (cudaGraphLaunch(h->run_graph, s) != cudaSuccess) { cudaGetLastError(); drop_run_graph(h); h->run_graph_failed = true; return false; }
    h->res_on_host = true;
    return true;
}

bool use_batched(const Handle *h) {
    return h->batch > 1 && h->batched_ok && h->cfg.batch_mode == 0 && h->d_xT != nullptr &&
           batched_smem_bytes(h->cols) <= batched_smem_bytes((uint32_t)h->cfg.max_cols);
}

// 4 + 4 bytes per non-zero in fp32, 2 + 4 in the half-precision mode, + row_ptr
uint64_t algorithmic_matrix_bytes(const Handle *h) {
    return h->nnz * (half_mode(h) ? 6ull : 8ull) + (h->rows + 1) * (h->nnz > 0xFFFFFFFFull ? 8ull : 4ull);
}

int launch_float(Handle *h, uint32_t k, cudaStream_t s, bool profile = false, bool to_host = false) {
    if (!h->have_matrix) return h->fail(TKS_ESTATE, "no matrix uploaded");
    if (!h->have_query) return h->fail(TKS_ESTATE, "no query set");
    if (k == 0 || k > h->kmax) return h->fail(TKS_EINVAL, "k=%u outside 1..%u", k, h->kmax);
    int rcd = pipe_drain(h);
    if (rcd) return rcd;
    h->last_run_pipelined = false;
    const uint64_t matrix_bytes = algorithmic_matrix_bytes(h);
    if (use_batched(h)) {
        launch_batched(h, k, s, profile);
        h->last_run_batched = true;
        h->stats.launches_per_run = 5;
        // matrix bytes once + Q*C*4 + Q*k*8 (the kernel re-reads the matrix once per 32 queries;
        // that is not counted)
        h->stats.algorithmic_bytes = matrix_bytes + (uint64_t)h->batch * ((uint64_t)h->cols * 4ull + k * 8ull);
    } else {
        // blocking tks_run of one query: one graph launch instead of three kernel launches
        const bool graphed = to_host && h->batch == 1 && !profile && s == h->stream && launch_query_graph(h, k, s);
        if (!graphed)
            for (uint32_t q = 0; q < h->batch; q++) launch_single_query(h, q, k, s, profile && q == 0, to_host && h->batch == 1);
        h->last_run_batched = false;
        h->stats.launches_per_run = 3 * h->batch;
        h->stats.algorithmic_bytes = matrix_bytes + (uint64_t)h->cols * 4ull + k * 8ull;
    }
    TKS_CUDA(h, cudaGetLastError());
    h->last_k = k;
    return TKS_OK;
}

// Results of the last run -> pinned host block (asynchronous on s).
int fetch_results_async(Handle *h, cudaStream_t s) {
    if (h->batch == (uint32_t)h->cfg.max_batch || h->res_block_bytes <= (64u << 10)) {
        TKS_CUDA(h, cudaMemcpyAsync(h->h_res_block, h->d_res_block, h->res_block_bytes, cudaMemcpyDeviceToHost, s));
    } else {
        const size_t n = (size_t)h->batch * h->kmax;
        TKS_CUDA(h, cudaMemcpyAsync(h->h_res_idx, h->d_res_idx, n * 4, cudaMemcpyDeviceToHost, s));
        TKS_CUDA(h, cudaMemcpyAsync(h->h_res_val, h->d_res_val, n * 4, cudaMemcpyDeviceToHost, s));
        TKS_CUDA(h, cudaMemcpyAsync(h->h_res_count, h->d_res_count, h->batch * 4, cudaMemcpyDeviceToHost, s));
    }
    return TKS_OK;
}

// Batched mode: queries whose candidate pool overflowed are re-run one by one through the single-query
// kernels (their per-warp buffers compact instead of overflowing).  Called with the results in the pinned
// host buffers; returns with them complete.
int resolve_batched_overflow(Handle *h, cudaStream_t s) {
    bool any = false;
    for (uint32_t q = 0; q < h->batch; q++) {
        if (h->h_res_count[q] != kPoolOverflow) continue;
        launch_single_query(h, q, h->last_k, s, false);
        any = true;
    }
    if (!any) return TKS_OK;
    TKS_CUDA(h, cudaGetLastError());
    int rc = fetch_results_async(h, s);
    if (rc) return rc;
    TKS_CUDA(h, cudaStreamSynchronize(s));
    h->stats.batched_fallbacks += 1;
    return TKS_OK;
}

}  // namespace

// Exclusive scan of n u32 values into n+1 u64 values (setup only).
int tks::device_scan_u32(Handle *h, const uint32_t *d_in, uint64_t n, uint64_t *d_out) {
    cudaStream_t s = h->stream;
    const uint32_t nb = (uint32_t)((n + kScanBlock - 1) / kScanBlock);
    uint64_t *d_bs = nullptr;
    TKS_CUDA(h, cudaMalloc(&d_bs, ((size_t)nb + 1) * sizeof(uint64_t)));
    scan_block_sums_kernel<<<nb, kScanBlock, 0, s>>>(d_in, n, d_bs);
    scan_block_offsets_kernel<<<1, 32, 0, s>>>(d_bs, nb);
    scan_finish_kernel<<<nb, kScanBlock, 0, s>>>(d_in, n, d_bs, d_out);
    TKS_CUDA(h, cudaStreamSynchronize(s));
    cudaFree(d_bs);
    return TKS_OK;
}

// ---------------------------------------------------------------------------

extern "C" {

int tks_version(void) { return TKS_VERSION; }

int tks_default_config(tks_config *cfg) {
    if (!cfg) return TKS_EINVAL;
    std::memset(cfg, 0, sizeof *cfg);
    cfg->mode = TKS_MODE_FLOAT_CSR;
    cfg->fixed_width = 20;            // the paper's headline design (BASELINE config 3); types.hpp:20 ships 32
    cfg->partitions = 32;             // types.hpp:36
    cfg->local_k = 8;                 // types.hpp:51
    cfg->limited_finished_rows = 4;   // types.hpp:77
    cfg->max_cols = 1024;             // types.hpp:55
    cfg->tie_break = TKS_TIE_LOWER_INDEX;
    cfg->device = 0;
    cfg->max_batch = 1;
    cfg->chunk_nnz = 0;
    return TKS_OK;
}

const char *tks_last_error(const tks_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int tks_create(const tks_config *cfg, tks_handle **out) {
    if (!cfg || !out) { g_create_error = "null argument"; return TKS_EINVAL; }
    *out = nullptr;
    if (cfg->mode != TKS_MODE_FLOAT_CSR && cfg->mode != TKS_MODE_FIXED_BSCSR) {
        g_create_error = "unknown mode"; return TKS_EINVAL;
    }
    if (cfg->max_cols < 1 || cfg->max_cols >= (int)kMaxCols) {
        g_create_error = "max_cols outside 1..16383"; return TKS_EINVAL;
    }
    if (cfg->mode == TKS_MODE_FIXED_BSCSR) {
        if (cfg->fixed_width < 17 || cfg->fixed_width > 32) { g_create_error = "fixed_width outside 17..32"; return TKS_EINVAL; }
        if (cfg->max_cols > 1024) { g_create_error = "BS-CSR column field is 10 bits: max_cols <= 1024"; return TKS_EINVAL; }
        const int B = tks_bscsr_packet_size(cfg->fixed_width);
        // the kernels are instantiated for LIMITED_FINISHED_ROWS 1..4 (types.hpp:77 ships 4; its "clean" design
        // LFR = BSCSR_PACKET_SIZE, types.hpp:76, is not built): say so at create, not at the first run
        if (cfg->limited_finished_rows < 1 || cfg->limited_finished_rows > B || cfg->limited_finished_rows > 4) {
            g_create_error = "limited_finished_rows outside 1..4"; return TKS_EINVAL;
        }
        if (cfg->local_k < 1 || cfg->local_k > 64) { g_create_error = "local_k outside 1..64"; return TKS_EINVAL; }
        if (cfg->partitions < 1 || cfg->partitions > 4096) { g_create_error = "partitions outside 1..4096"; return TKS_EINVAL; }
    }
    if (cfg->value_type != TKS_VALUE_FP32 && cfg->value_type != TKS_VALUE_FP16 && cfg->value_type != TKS_VALUE_BF16) {
        g_create_error = "unknown value_type"; return TKS_EINVAL;
    }
    if (cfg->value_type != TKS_VALUE_FP32 && cfg->mode != TKS_MODE_FLOAT_CSR) {
        g_create_error = "16-bit value types belong to FLOAT_CSR mode"; return TKS_EINVAL;
    }
    if (cfg->max_batch < 1 || cfg->max_batch > 1024) { g_create_error = "max_batch outside 1..1024"; return TKS_EINVAL; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e);
        return TKS_ECUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return TKS_EINVAL; }
    tks_handle *h = new (std::nothrow) tks_handle();
    if (!h) { g_create_error = "out of memory"; return TKS_ENOMEM; }
    h->cfg = *cfg;
    h->device = cfg->device;
    auto bail = [&](const char *what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        tks_destroy(h);
        return TKS_ECUDA;
    };
    if ((e = cudaSetDevice(h->device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major < 10) {
        g_create_error = "device is not sm_100-class (this library is built for sm_100a only)";
        tks_destroy(h);
        return TKS_ECUDA;
    }
    h->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&h->ev_query, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->evm0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->evm1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if (cfg->mode == TKS_MODE_FLOAT_CSR) {
        if ((e = prep_main<256>(h, 0)) != cudaSuccess) return bail("prep_main<256>", e);
        if ((e = prep_main<512>(h, 1)) != cudaSuccess) return bail("prep_main<512>", e);
        if ((e = prep_main<1024>(h, 2)) != cudaSuccess) return bail("prep_main<1024>", e);
        if ((e = prep_main<2048>(h, 3)) != cudaSuccess) return bail("prep_main<2048>", e);
        {
            size_t ss = ((size_t)cfg->max_cols + 1u) * 4u;
            if (ss < (size_t)kHistScratchWords * 4u) ss = (size_t)kHistScratchWords * 4u;
            if ((e = cudaFuncSetAttribute(csr_sample_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(csr_sample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(csr_sample_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss)) != cudaSuccess)
                return bail("sample smem attr", e);
            if ((e = cudaFuncSetAttribute(select_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(kSelectDynSmem))) != cudaSuccess ||
                (e = cudaFuncSetAttribute(select_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(kSelectDynSmem))) != cudaSuccess)
                return bail("select smem attr", e);
            const int mx = carveout_pct();
            if (mx >= 0 && ((e = cudaFuncSetAttribute(csr_sample_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, mx)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(csr_sample_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, mx)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(csr_sample_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, mx)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(select_topk_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(select_topk_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, mx)) != cudaSuccess))
                return bail("carve-out attr", e);
        }
        if (cfg->max_batch > 1 && batched_smem_bytes((uint32_t)cfg->max_cols) <= (size_t)prop.sharedMemPerBlockOptin) {
            const int bs = (int)batched_smem_bytes((uint32_t)cfg->max_cols);
            if ((e = cudaFuncSetAttribute(csr_batched_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bs)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(csr_batched_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bs)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(csr_batched_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bs)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(csr_batched_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bs)) != cudaSuccess)
                return bail("batched smem attr", e);
            h->batched_ok = true;
        }
        int rc = alloc_query_side(h);
        if (rc != TKS_OK) { g_create_error = h->err; tks_destroy(h); return rc; }
    }
    *out = h;
    return TKS_OK;
}

void tks_destroy(tks_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    free_matrix(h);
    bscsr_destroy(h);
    for (uint32_t r = 0; r < h->peer_world; r++)
        if (r != h->peer_rank && h->peer_mapped[r]) cudaIpcCloseMemHandle(h->peer_mapped[r]);
    cudaFree(h->d_peer_window);
    cudaFree(h->d_x); cudaFree(h->d_state); cudaFree(h->d_pool); cudaFree(h->d_sample_keys);
    cudaFree(h->d_res_keys); cudaFree(h->d_res_block);
    cudaFree(h->d_xT); cudaFree(h->d_bpool); cudaFree(h->d_pass_counter); cudaFree(h->d_bsample_keys);
    cudaFreeHost(h->h_res_block); cudaFreeHost(h->h_x);
    drop_run_graph(h);
    cudaFree(h->d_pipe_state); cudaFree(h->d_pipe_sample_keys); cudaFree(h->d_pipe_stamps);
    cudaFreeHost(h->h_pipe_stamps);
    cudaFree(h->d_pipe_query); cudaFreeHost(h->h_pipe_query); cudaFreeHost(h->h_pipe_res);
    for (int i = 0; i < tks::Handle::kPipeSlots; i++) {
        cudaFree(h->d_pipe_pool[i]);
        if (h->pipe_ev_done[i]) cudaEventDestroy(h->pipe_ev_done[i]);
    }
    if (h->pipe_ev_query) cudaEventDestroy(h->pipe_ev_query);
    if (h->pipe_sample_stream) cudaStreamDestroy(h->pipe_sample_stream);
    if (h->pipe_select_stream) cudaStreamDestroy(h->pipe_select_stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_query) cudaEventDestroy(h->ev_query);
    if (h->evm0) cudaEventDestroy(h->evm0);
    if (h->evm1) cudaEventDestroy(h->evm1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int tks_upload_csr_device(tks_handle *h, uint64_t rows, uint32_t cols, uint64_t nnz, const void *d_ptr,
                          int ptr_bits, const uint32_t *d_idx, const float *d_val, uint64_t row_offset) {
    if (!h) return TKS_EINVAL;
    TKS_CUDA(h, cudaSetDevice(h->device));
    int rc = check_float_upload(h, rows, cols, nnz, row_offset);
    if (rc) return rc;
    if (ptr_bits != 32 && ptr_bits != 64) return h->fail(TKS_EINVAL, "ptr_bits must be 32 or 64");
    if (!d_ptr || (nnz && (!d_idx || !d_val))) return h->fail(TKS_EINVAL, "null array");
    free_matrix(h);
    h->row_offset = row_offset;
    // keep a 64-bit copy of row_ptr for tks_download_csr
    TKS_CUDA(h, cudaMalloc(&h->d_ptr64, (rows + 1) * sizeof(uint64_t)));
    if (ptr_bits == 64) {
        TKS_CUDA(h, cudaMemcpyAsync(h->d_ptr64, d_ptr, (rows + 1) * 8, cudaMemcpyDeviceToDevice, h->stream));
        return build_from_device_csr<uint64_t>(h, rows, cols, nnz, (const uint64_t *)d_ptr, d_idx, d_val, nullptr);
    }
    int rc2 = build_from_device_csr<uint32_t>(h, rows, cols, nnz, (const uint32_t *)d_ptr, d_idx, d_val, nullptr);
    if (rc2) return rc2;
    widen_u32_to_u64_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>((const uint32_t *)d_ptr, rows + 1, h->d_ptr64);
    TKS_CUDA(h, cudaStreamSynchronize(h->stream));
    return TKS_OK;
}

int tks_upload_csr(tks_handle *h, uint64_t rows, uint32_t cols, uint64_t nnz, const void *ptr, int ptr_bits,
                   const uint32_t *idx, const float *val, uint64_t row_offset) {
    if (!h) return TKS_EINVAL;
    TKS_CUDA(h, cudaSetDevice(h->device));
    int rc = check_float_upload(h, rows, cols, nnz, row_offset);
    if (rc) return rc;
    if (ptr_bits != 32 && ptr_bits != 64) return h->fail(TKS_EINVAL, "ptr_bits must be 32 or 64");
    if (!ptr || (nnz && (!idx || !val))) return h->fail(TKS_EINVAL, "null array");
    void *d_ptr = nullptr;
    uint32_t *d_idx = nullptr;
    float *d_val = nullptr;
    const size_t pb = (size_t)(ptr_bits / 8);
    TKS_CUDA(h, cudaMalloc(&d_ptr, (rows + 1) * pb));
    TKS_CUDA(h, cudaMalloc(&d_idx, (nnz ? nnz : 1) * sizeof(uint32_t)));
    TKS_CUDA(h, cudaMalloc(&d_val, (nnz ? nnz : 1) * sizeof(float)));
    TKS_CUDA(h, cudaMemcpy(d_ptr, ptr, (rows + 1) * pb, cudaMemcpyHostToDevice));
    if (nnz) {
        TKS_CUDA(h, cudaMemcpy(d_idx, idx, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice));
        TKS_CUDA(h, cudaMemcpy(d_val, val, nnz * sizeof(float), cudaMemcpyHostToDevice));
    }
    rc = tks_upload_csr_device(h, rows, cols, nnz, d_ptr, ptr_bits, d_idx, d_val, row_offset);
    cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_val);
    return rc;
}

int tks_generate_synthetic(tks_handle *h, uint64_t rows, uint32_t cols, uint32_t avg_degree, int dist,
                           uint64_t seed, uint64_t row_offset) {
    if (!h) return TKS_EINVAL;
    TKS_CUDA(h, cudaSetDevice(h->device));
    int rc = check_float_upload(h, rows, cols, 0, row_offset);
    if (rc) return rc;
    if (rows == 0 || avg_degree < 2 || avg_degree > 128) return h->fail(TKS_EINVAL, "rows>0, 2<=avg_degree<=128");
    if (dist != 0 && dist != 1) return h->fail(TKS_EINVAL, "dist: 0 uniform, 1 gamma");
    free_matrix(h);
    h->row_offset = row_offset;
    cudaStream_t s = h->stream;
    uint32_t *d_deg = nullptr;
    TKS_CUDA(h, cudaMalloc(&d_deg, rows * sizeof(uint32_t)));
    TKS_CUDA(h, cudaMalloc(&h->d_ptr64, (rows + 1) * sizeof(uint64_t)));
    synth_degree_kernel<<<(uint32_t)((rows + 255) / 256), 256, 0, s>>>(rows, row_offset, seed, avg_degree, dist, d_deg);
    rc = device_scan_u32(h, d_deg, rows, h->d_ptr64);
    if (rc) return rc;
    uint64_t nnz = 0;
    TKS_CUDA(h, cudaMemcpyAsync(&nnz, h->d_ptr64 + rows, 8, cudaMemcpyDeviceToHost, s));
    TKS_CUDA(h, cudaStreamSynchronize(s));
    cudaFree(d_deg);
    uint32_t *d_idx = nullptr;
    float *d_val = nullptr;
    const size_t pad = 2048;
    TKS_CUDA(h, cudaMalloc(&d_idx, nnz * sizeof(uint32_t)));
    TKS_CUDA(h, cudaMalloc(&d_val, nnz * sizeof(float) + pad));
    TKS_CUDA(h, cudaMemsetAsync(reinterpret_cast<uint8_t *>(d_val) + nnz * sizeof(float), 0, pad, s));
    synth_fill_kernel<<<(uint32_t)((rows + 127) / 128), 128, 0, s>>>(rows, row_offset, seed, cols, h->d_ptr64, d_idx, d_val);
    TKS_CUDA(h, cudaGetLastError());
    rc = build_from_device_csr<uint64_t>(h, rows, cols, nnz, h->d_ptr64, d_idx, nullptr, d_val);
    cudaFree(d_idx);
    return rc;
}

int tks_download_csr_rows(tks_handle *h, uint64_t row_begin, uint64_t row_end, uint64_t *ptr64, uint32_t *idx, float *val) {
    if (!h) return TKS_EINVAL;
    if (!h->have_matrix || !h->d_ptr64) return h->fail(TKS_ESTATE, "no CSR matrix resident");
    if (row_begin > row_end || row_end > h->rows) return h->fail(TKS_EINVAL, "row range outside 0..rows");
    TKS_CUDA(h, cudaSetDevice(h->device));
    uint64_t ends[2] = {0, 0};
    TKS_CUDA(h, cudaMemcpy(&ends[0], h->d_ptr64 + row_begin, 8, cudaMemcpyDeviceToHost));
    TKS_CUDA(h, cudaMemcpy(&ends[1], h->d_ptr64 + row_end, 8, cudaMemcpyDeviceToHost));
    const uint64_t b = ends[0], n = ends[1] - ends[0];
    if (ptr64) TKS_CUDA(h, cudaMemcpy(ptr64, h->d_ptr64 + row_begin, (row_end - row_begin + 1) * 8, cudaMemcpyDeviceToHost));
    if (val && !half_mode(h)) TKS_CUDA(h, cudaMemcpy(val, reinterpret_cast<const float *>(h->d_val) + b, n * 4, cudaMemcpyDeviceToHost));
    if (val && half_mode(h)) {
        // the resident values are halves: fetch them into the upper half of the output and widen in place
        uint16_t *tmp = reinterpret_cast<uint16_t *>(val) + n;
        TKS_CUDA(h, cudaMemcpy(tmp, reinterpret_cast<const uint16_t *>(h->d_val) + b, n * 2, cudaMemcpyDeviceToHost));
        if (value_type(h) == TKS_VALUE_BF16) {
            for (uint64_t i = 0; i < n; i++) { const uint32_t w = (uint32_t)tmp[i] << 16; std::memcpy(&val[i], &w, 4); }
        } else {
            for (uint64_t i = 0; i < n; i++) val[i] = half_bits_to_float(tmp[i]);
        }
    }
    if (idx) {
        // col16 holds column * 4: fetch the 16-bit words into the upper half of the output, widen in place
        uint16_t *tmp = reinterpret_cast<uint16_t *>(idx) + n;
        TKS_CUDA(h, cudaMemcpy(tmp, h->d_col16 + b, n * 2, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n; i++) idx[i] = (uint32_t)tmp[i] >> 2;
    }
    return TKS_OK;
}

int tks_download_csr(tks_handle *h, uint64_t *ptr64, uint32_t *idx, float *val) {
    if (!h) return TKS_EINVAL;
    return tks_download_csr_rows(h, 0, h->rows, ptr64, idx, val);
}

int tks_set_query_device(tks_handle *h, const void *d_vec, uint32_t batch, void *cuda_stream) {
    if (!h || !d_vec) return TKS_EINVAL;
    TKS_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) {
        if (batch != 1) return h->fail(TKS_EINVAL, "BS-CSR mode takes one query at a time");
        return bscsr_set_query(h, nullptr, (const uint32_t *)d_vec, s);
    }
    if (!h->have_matrix) return h->fail(TKS_ESTATE, "no matrix uploaded");
    if (batch < 1 || batch > (uint32_t)h->cfg.max_batch) return h->fail(TKS_EINVAL, "batch outside 1..max_batch");
    TKS_CUDA(h, cudaMemcpyAsync(h->d_x, d_vec, (size_t)batch * h->cols * sizeof(float), cudaMemcpyDeviceToDevice, s));
    h->batch = batch;
    h->have_query = true;
    return TKS_OK;
}

int tks_set_query(tks_handle *h, const void *vec, uint32_t batch) {
    if (!h || !vec) return TKS_EINVAL;
    TKS_CUDA(h, cudaSetDevice(h->device));
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) {
        if (batch != 1) return h->fail(TKS_EINVAL, "BS-CSR mode takes one query at a time");
        return bscsr_set_query(h, (const uint32_t *)vec, nullptr, h->stream);
    }
    if (!h->have_matrix) return h->fail(TKS_ESTATE, "no matrix uploaded");
    if (batch < 1 || batch > (uint32_t)h->cfg.max_batch) return h->fail(TKS_EINVAL, "batch outside 1..max_batch");
    const size_t bytes = (size_t)batch * h->cols * sizeof(float);
    TKS_CUDA(h, cudaEventSynchronize(h->ev_query));   // the pinned staging buffer may still feed the previous copy
    std::memcpy(h->h_x, vec, bytes);
    TKS_CUDA(h, cudaMemcpyAsync(h->d_x, h->h_x, bytes, cudaMemcpyHostToDevice, h->stream));
    TKS_CUDA(h, cudaEventRecord(h->ev_query, h->stream));
    h->batch = batch;
    h->have_query = true;
    return TKS_OK;
}

int tks_run_async(tks_handle *h, uint32_t k, void *cuda_stream) {
    if (!h) return TKS_EINVAL;
    TKS_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) { h->last_k = k; return bscsr_launch(h, s); }
    // a query staged by tks_set_query (host) travels on the handle's own stream
    if (s != h->stream) TKS_CUDA(h, cudaStreamWaitEvent(s, h->ev_query, 0));
    int rc = launch_float(h, k, s);
    if (rc) return rc;
    h->have_result = false;
    h->overflow_check_pending = h->last_run_batched;
    return TKS_OK;
}

int tks_run(tks_handle *h, uint32_t k, float *kernel_ms, float *total_ms) {
    if (!h) return TKS_EINVAL;
    TKS_CUDA(h, cudaSetDevice(h->device));
    auto t0 = std::chrono::high_resolution_clock::now();
    TKS_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    int rc;
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) { h->last_k = k; rc = bscsr_launch(h, h->stream); }
    else rc = launch_float(h, k, h->stream, h->cfg.profile_kernels != 0, true);
    if (rc) return rc;
    TKS_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) {
        rc = bscsr_fetch(h);
        if (rc) return rc;
    } else {
        const bool direct = !h->last_run_batched && h->batch == 1;   // the select kernel wrote the pinned host block itself
        if (!direct) {
            rc = fetch_results_async(h, h->stream);
            if (rc) return rc;
        }
        TKS_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h->last_run_batched) {
            rc = resolve_batched_overflow(h, h->stream);
            if (rc) return rc;
        }
        h->overflow_check_pending = false;
    }
    float ms = 0.f;
    TKS_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    auto t1 = std::chrono::high_resolution_clock::now();
    h->stats.last_kernel_ms = ms;
    if (h->cfg.profile_kernels) {
        float mm = 0.f;
        TKS_CUDA(h, cudaEventElapsedTime(&mm, h->evm0, h->evm1));
        h->stats.last_main_kernel_ms = mm;
    }
    h->stats.last_total_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
    if (kernel_ms) *kernel_ms = ms;
    if (total_ms) *total_ms = h->stats.last_total_ms;
    h->have_result = true;
    return TKS_OK;
}

int tks_read_result(tks_handle *h, uint32_t query, uint32_t *idx_out, void *val_out, uint32_t *count) {
    if (!h || !idx_out || !val_out) return TKS_EINVAL;
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) {
        if (query != 0) return h->fail(TKS_EINVAL, "BS-CSR mode has a single query");
        return bscsr_read_result(h, idx_out, (uint32_t *)val_out, h->last_k, count);
    }
    if (h->last_run_pipelined && h->res_on_host)
        return h->fail(TKS_ESTATE, "the last query went through tks_submit_host: its result is read with tks_fetch(ticket)");
    if (!h->have_result) {
        // results of an async run: fetch now
        if (h->last_k == 0) return h->fail(TKS_ESTATE, "no run yet");
        TKS_CUDA(h, cudaSetDevice(h->device));
        TKS_CUDA(h, cudaDeviceSynchronize());
        int rcd = pipe_drain(h);
        if (rcd) return rcd;
        int rcf = fetch_results_async(h, h->stream);
        if (rcf) return rcf;
        TKS_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h->overflow_check_pending) {
            int rc = resolve_batched_overflow(h, h->stream);
            if (rc) return rc;
            h->overflow_check_pending = false;
        }
        h->have_result = true;
    }
    if (query >= h->batch) return h->fail(TKS_EINVAL, "query index out of range");
    if (h->h_res_count[query] == kPeerTimeout)
        return h->fail(TKS_ECUDA, "device-side wait timed out (TKS_SPIN_TIMEOUT_MS): a rank of the box never delivered its "
                                  "candidates, or the main kernel of a pipelined submit never completed");
    const uint32_t k = h->last_k;
    std::memcpy(idx_out, h->h_res_idx + (size_t)query * h->kmax, k * 4);
    std::memcpy(val_out, h->h_res_val + (size_t)query * h->kmax, k * 4);
    if (count) *count = h->h_res_count[query];
    return TKS_OK;
}

int tks_read_partition_results(tks_handle *h, uint32_t *idx_words, uint32_t *val_words) {
    if (!h) return TKS_EINVAL;
    if (h->cfg.mode != TKS_MODE_FIXED_BSCSR) return h->fail(TKS_ESTATE, "BS-CSR mode only");
    return bscsr_read_partition_results(h, idx_words, val_words);
}

int tks_partition_words_device(tks_handle *h, const uint32_t **d_words, uint32_t *n_words) {
    if (!h || !d_words) return TKS_EINVAL;
    if (h->cfg.mode != TKS_MODE_FIXED_BSCSR) return h->fail(TKS_ESTATE, "BS-CSR mode only");
    return bscsr_partition_words_device(h, d_words, n_words);
}

int tks_result_keys_device(tks_handle *h, uint32_t query, const uint64_t **d_keys, uint32_t *count) {
    if (!h || !d_keys) return TKS_EINVAL;
    if (h->cfg.mode != TKS_MODE_FLOAT_CSR) return h->fail(TKS_ESTATE, "float mode only");
    if (query >= (uint32_t)h->cfg.max_batch) return h->fail(TKS_EINVAL, "query index out of range");
    if (h->overflow_check_pending) {
        // an async batched run: make sure no query's candidate pool overflowed before its keys are used
        TKS_CUDA(h, cudaSetDevice(h->device));
        TKS_CUDA(h, cudaDeviceSynchronize());
        TKS_CUDA(h, cudaMemcpy(h->h_res_count, h->d_res_count, h->batch * 4, cudaMemcpyDeviceToHost));
        bool any = false;
        for (uint32_t q = 0; q < h->batch; q++) any |= h->h_res_count[q] == kPoolOverflow;
        if (any) {
            int rc = resolve_batched_overflow(h, h->stream);
            if (rc) return rc;
        }
        h->overflow_check_pending = false;
    }
    *d_keys = h->d_res_keys + (size_t)query * h->kmax;
    if (count) *count = h->last_k;
    return TKS_OK;
}

int tks_merge_keys_device(tks_handle *h, uint32_t query, const uint64_t *d_keys, uint32_t n_keys, uint32_t k,
                          void *cuda_stream) {
    if (!h || !d_keys) return TKS_EINVAL;
    if (h->cfg.mode != TKS_MODE_FLOAT_CSR) return h->fail(TKS_ESTATE, "float mode only");
    if (query >= (uint32_t)h->cfg.max_batch) return h->fail(TKS_EINVAL, "query index out of range");
    if (k == 0 || k > h->kmax) return h->fail(TKS_EINVAL, "k out of range");
    TKS_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    select_topk_kernel<false><<<1, kSelectThreads, kSelectDynSmem, s>>>(
        d_keys, 0u, nullptr, n_keys, 0u, k, h->cfg.tie_break == TKS_TIE_HIGHER_INDEX,
        h->d_res_keys + (size_t)query * h->kmax, h->d_res_idx + (size_t)query * h->kmax,
        h->d_res_val + (size_t)query * h->kmax, 0u, h->d_res_count + query, nullptr, PeerExchange{}, 0u, 0u, 0u, nullptr, kSelectSmemKeys);
    h->res_on_host = false;
    TKS_CUDA(h, cudaGetLastError());
    h->last_k = k;
    if (h->batch < query + 1) h->batch = query + 1;
    h->have_result = false;   // tks_read_result will fetch
    return TKS_OK;
}

int tks_merge_keys_batched_device(tks_handle *h, const uint64_t *d_keys, uint32_t keys_per_query, uint32_t batch,
                                  uint32_t k, void *cuda_stream) {
    if (!h || !d_keys) return TKS_EINVAL;
    if (h->cfg.mode != TKS_MODE_FLOAT_CSR) return h->fail(TKS_ESTATE, "float mode only");
    if (batch < 1 || batch > (uint32_t)h->cfg.max_batch) return h->fail(TKS_EINVAL, "batch outside 1..max_batch");
    if (k == 0 || k > h->kmax) return h->fail(TKS_EINVAL, "k out of range");
    TKS_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    select_topk_kernel<false><<<batch, kSelectThreads, kSelectDynSmem, s>>>(
        d_keys, keys_per_query, nullptr, keys_per_query, 0u, k, h->cfg.tie_break == TKS_TIE_HIGHER_INDEX,
        h->d_res_keys, h->d_res_idx, h->d_res_val, h->kmax, h->d_res_count, nullptr, PeerExchange{}, 0u, 0u, 0u, nullptr, kSelectSmemKeys);
    h->res_on_host = false;
    TKS_CUDA(h, cudaGetLastError());
    h->last_k = k;
    if (h->batch < batch) h->batch = batch;
    h->have_result = false;
    h->overflow_check_pending = false;
    return TKS_OK;
}

// ---- peer-memory candidate exchange ------------------------------------------------------------------------------

int tks_peer_init(tks_handle *h, uint32_t world, uint32_t rank, void *ipc_handle_out) {
    if (!h || !ipc_handle_out) return TKS_EINVAL;
    if (h->cfg.mode != TKS_MODE_FLOAT_CSR) return h->fail(TKS_ESTATE, "float mode only");
    if (world < 1 || world > kPeerMaxWorld || rank >= world) return h->fail(TKS_EINVAL, "world outside 1..8 or rank >= world");
    static_assert(sizeof(cudaIpcMemHandle_t) <= TKS_IPC_HANDLE_BYTES, "IPC handle does not fit");
    TKS_CUDA(h, cudaSetDevice(h->device));
    {
        // the step counter restarts at 0 below: records of an earlier session must not satisfy the new steps' polls
        int rcd = pipe_drain(h);
        if (rcd) return rcd;
        const size_t bytes = peer_window_bytes(h->kmax);
        if (!h->d_peer_window) TKS_CUDA(h, cudaMalloc(&h->d_peer_window, bytes));
        TKS_CUDA(h, cudaDeviceSynchronize());
        TKS_CUDA(h, cudaMemset(h->d_peer_window, 0, bytes));
    }
    cudaIpcMemHandle_t ih;
    TKS_CUDA(h, cudaIpcGetMemHandle(&ih, h->d_peer_window));
    std::memset(ipc_handle_out, 0, TKS_IPC_HANDLE_BYTES);
    std::memcpy(ipc_handle_out, &ih, sizeof ih);
    h->peer_world = world; h->peer_rank = rank; h->peer_seq = 0; h->peer_ready = false;
    return TKS_OK;
}

int tks_peer_connect(tks_handle *h, const void *all_handles) {
    if (!h || !all_handles) return TKS_EINVAL;
    if (!h->d_peer_window || h->peer_world == 0) return h->fail(TKS_ESTATE, "tks_peer_init first");
    TKS_CUDA(h, cudaSetDevice(h->device));
    for (uint32_t r = 0; r < h->peer_world; r++) {
        if (r == h->peer_rank) { h->peer_mapped[r] = h->d_peer_window; continue; }
        cudaIpcMemHandle_t ih;
        std::memcpy(&ih, static_cast<const uint8_t *>(all_handles) + (size_t)r * TKS_IPC_HANDLE_BYTES, sizeof ih);
        TKS_CUDA(h, cudaIpcOpenMemHandle(&h->peer_mapped[r], ih, cudaIpcMemLazyEnablePeerAccess));
    }
    h->peer_ready = true;
    return TKS_OK;
}

static PeerExchange peer_args(const tks_handle *h) {
    PeerExchange px{};
    for (uint32_t r = 0; r < h->peer_world; r++) px.window[r] = static_cast<uint64_t *>(h->peer_mapped[r]);
    px.world = h->peer_world; px.rank = h->peer_rank; px.kmax = h->kmax;
    px.timeout_ms = spin_timeout_ms();
    return px;
}

int tks_run_exchange_async(tks_handle *h, uint32_t k, void *cuda_stream) {
    if (!h) return TKS_EINVAL;
    if (!h->peer_ready) return h->fail(TKS_ESTATE, "tks_peer_connect first");
    if (h->cfg.mode != TKS_MODE_FLOAT_CSR) return h->fail(TKS_ESTATE, "float mode only");
    if (!h->have_matrix) return h->fail(TKS_ESTATE, "no matrix uploaded");
    if (!h->have_query) return h->fail(TKS_ESTATE, "no query set");
    if (h->batch != 1) return h->fail(TKS_EINVAL, "the peer exchange serves one query per run (batched runs use the all-gather path)");
    if (k == 0 || k > h->kmax) return h->fail(TKS_EINVAL, "k=%u outside 1..%u", k, h->kmax);
    if ((uint64_t)h->peer_world * k > kSelectSortCap) return h->fail(TKS_EINVAL, "world * k exceeds %u", kSelectSortCap);
    TKS_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    if (s != h->stream) TKS_CUDA(h, cudaStreamWaitEvent(s, h->ev_query, 0));
    int rcd = pipe_drain(h);
    if (rcd) return rcd;
    const PeerExchange px = peer_args(h);
    h->peer_seq += 1;
    // sample -> main -> select; the select kernel also exchanges the candidates with the peers and merges
    launch_single_query(h, 0, k, s, false, false, &px, h->peer_seq);
    TKS_CUDA(h, cudaGetLastError());
    h->last_run_batched = false;
    h->last_run_pipelined = false;
    h->stats.launches_per_run = 3;
    h->stats.algorithmic_bytes = algorithmic_matrix_bytes(h) + (uint64_t)h->cols * 4ull + k * 8ull;
    h->last_k = k;
    h->have_result = false;
    h->overflow_check_pending = false;
    return TKS_OK;
}

int tks_peer_exchange_async(tks_handle *h, uint32_t k, void *cuda_stream) {
    if (!h) return TKS_EINVAL;
    if (!h->peer_ready) return h->fail(TKS_ESTATE, "tks_peer_connect first");
    if (k != h->last_k || h->batch != 1) return h->fail(TKS_ESTATE, "no single-query run with this k precedes the exchange");
    if (h->res_on_host || h->last_run_pipelined)
        return h->fail(TKS_ESTATE, "the last run left its result in host memory (blocking tks_run) or was a pipelined submit: "
                                   "run with tks_run_async before tks_peer_exchange_async");
    if ((uint64_t)h->peer_world * k > kSelectSortCap) return h->fail(TKS_EINVAL, "world * k exceeds %u", kSelectSortCap);
    TKS_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    const PeerExchange px = peer_args(h);
    h->peer_seq += 1;
    TKS_CUDA(h, launch_pdl(peer_exchange_merge_kernel, dim3(1), dim3(kSelectThreads), (size_t)0, s, pdl_enabled(), px,
                           h->peer_seq, k, (int)(h->cfg.tie_break == TKS_TIE_HIGHER_INDEX), h->d_res_keys, h->d_res_idx,
                           h->d_res_val, h->d_res_count));
    return TKS_OK;
}

// ---- pipelined submits ---------------------------------------------------------------------------------------------

}  // extern "C"

namespace {

size_t pipe_res_words(const Handle *h) { return 64u + 2u * (size_t)h->kmax; }

// One query into the pipeline.  host_query != nullptr: the query comes from host memory (copied on the sample stream
// in front of the sample kernel) and the result goes to the slot's pinned host block (tks_fetch).
int pipe_submit(tks_handle *h, const float *d_query, const float *host_query, uint32_t k, uint32_t flags, cudaStream_t s,
                uint64_t *ticket) {
    if (h->cfg.mode != TKS_MODE_FLOAT_CSR) return h->fail(TKS_ESTATE, "float mode only");
    if (!h->have_matrix) return h->fail(TKS_ESTATE, "no matrix uploaded");
    if (k == 0 || k > h->kmax) return h->fail(TKS_EINVAL, "k=%u outside 1..%u", k, h->kmax);
    const bool exchange = (flags & TKS_SUBMIT_EXCHANGE) != 0;
    if (exchange) {
        if (!h->peer_ready) return h->fail(TKS_ESTATE, "tks_peer_connect first");
        if ((uint64_t)h->peer_world * k > kSelectSortCap) return h->fail(TKS_EINVAL, "world * k exceeds %u", kSelectSortCap);
    }
    TKS_CUDA(h, cudaSetDevice(h->device));
    int rc = pipe_init(h);
    if (rc) return rc;
    const uint32_t seq = h->pipe_seq + 1u;
    const int slot = (int)(seq % (uint32_t)h->pipe_slots);
    if (h->pipe_busy[slot]) {
        // the slot's previous query (pipe_slots submits ago) must have been selected before its scratch is reused: this
        // is the only place a submit blocks the host, and it bounds the queries in flight.  Four slots let the host
        // enqueue the sample of a query a whole step before its main kernel needs the threshold (the sample takes
        // ~130 us beside a running main kernel), and let the select of a step wait for a slow peer GPU for up to two
        // steps without stalling this GPU's stream.
        TKS_CUDA(h, cudaEventSynchronize(h->pipe_ev_done[slot]));
        h->pipe_busy[slot] = false;
    }
    const int variant = cap_variant_for_k(k);
    const int tie_higher = h->cfg.tie_break == TKS_TIE_HIGHER_INDEX;
    const CsrDevice m = csr_device(h);
    RunState *st = h->d_pipe_state + slot;
    uint32_t *o_idx = h->d_res_idx, *o_cnt = h->d_res_count;
    float *o_val = h->d_res_val;
    if (host_query) {
        float *hq = h->h_pipe_query + (size_t)slot * h->cfg.max_cols, *dq = h->d_pipe_query + (size_t)slot * h->cfg.max_cols;
        std::memcpy(hq, host_query, (size_t)h->cols * sizeof(float));
        TKS_CUDA(h, cudaMemcpyAsync(dq, hq, (size_t)h->cols * sizeof(float), cudaMemcpyHostToDevice, h->pipe_sample_stream));
        d_query = dq;
        // the select kernel stores indices, scores and the count straight into the slot's pinned host block
        uint32_t *blk = h->h_pipe_res + (size_t)slot * pipe_res_words(h);
        o_cnt = blk; o_idx = blk + 64; o_val = reinterpret_cast<float *>(blk + 64 + h->kmax);
        h->pipe_slot_ticket[slot] = seq;
    } else {
        h->pipe_slot_ticket[slot] = 0;
        if (!(flags & TKS_SUBMIT_QUERY_READY)) {
            // the query is produced by earlier work of the caller's stream: the sample stream has to see it too
            TKS_CUDA(h, cudaEventRecord(h->pipe_ev_query, s));
            TKS_CUDA(h, cudaStreamWaitEvent(h->pipe_sample_stream, h->pipe_ev_query, 0));
        }
    }
    h->pipe_slot_k[slot] = k;
    // 1. threshold of THIS query on the sample stream: small CTAs that fit beside the main kernel still streaming the
    //    previous query
    uint64_t *stamp = h->d_pipe_stamps + (size_t)(seq % tks::Handle::kPipeStamps) * kStampWords;
    launch_sample(h, m, d_query, st, h->d_pipe_sample_keys, k, h->pipe_sample_stream, (uint32_t)pipe_sample_threads(value_type(h)), seq, stamp);
    // 2. the stream on the caller's stream, chained to the previous main kernel by programmatic dependent launch and
    //    never waiting for it: its CTAs take over as that grid's CTAs retire
    launch_main_variant(h, variant, m, d_query, st, h->d_pipe_pool[slot], k, s, pdl_enabled(), seq, stamp);
    // 3. select (+ exchange over the peer windows + merge) on the select stream; waits for the main kernel's last CTA
    constexpr uint32_t lean_threads = kSelectLeanThreads;
    if (exchange && h->peer_world > 1) {
        const PeerExchange px = peer_args(h);
        h->peer_seq += 1;
        select_topk_kernel<true><<<1, lean_threads, lean_threads / 32u * 1024u, h->pipe_select_stream>>>(
            h->d_pipe_pool[slot], 0u, st, 0u, 0u, k, tie_higher, h->d_res_keys, o_idx, o_val, 0u,
            o_cnt, nullptr, px, h->peer_seq, seq, spin_timeout_ms(), stamp, 0u);
    } else {
        select_topk_kernel<false><<<1, lean_threads, lean_threads / 32u * 1024u, h->pipe_select_stream>>>(
            h->d_pipe_pool[slot], 0u, st, 0u, 0u, k, tie_higher, h->d_res_keys, o_idx, o_val, 0u,
            o_cnt, nullptr, PeerExchange{}, 0u, seq, spin_timeout_ms(), stamp, 0u);
    }
    TKS_CUDA(h, cudaGetLastError());
    TKS_CUDA(h, cudaEventRecord(h->pipe_ev_done[slot], h->pipe_select_stream));
    h->pipe_busy[slot] = true;
    h->pipe_seq = seq;
    h->pipe_last_slot = slot;
    h->batch = 1;
    h->last_k = k;
    h->have_result = false;
    h->res_on_host = host_query != nullptr;
    h->last_run_batched = false;
    h->last_run_pipelined = true;
    h->overflow_check_pending = false;
    h->stats.launches_per_run = 3;
    h->stats.algorithmic_bytes = algorithmic_matrix_bytes(h) + (uint64_t)h->cols * 4ull + k * 8ull;
    if (ticket) *ticket = seq;
    return TKS_OK;
}

}  // namespace

extern "C" {

int tks_submit(tks_handle *h, const void *d_query_v, uint32_t k, uint32_t flags, void *cuda_stream) {
    if (!h || !d_query_v) return TKS_EINVAL;
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) {
        if (flags & TKS_SUBMIT_EXCHANGE) return h->fail(TKS_EINVAL, "BS-CSR mode spreads partitions, not candidates (ShardedSpMVFixed)");
        TKS_CUDA(h, cudaSetDevice(h->device));
        return bscsr_submit(h, nullptr, static_cast<const uint32_t *>(d_query_v), k, cuda_stream ? (cudaStream_t)cuda_stream : h->stream,
                            (flags & TKS_SUBMIT_QUERY_READY) != 0, nullptr);
    }
    const float *d_query = static_cast<const float *>(d_query_v);
    return pipe_submit(h, d_query, nullptr, k, flags, cuda_stream ? (cudaStream_t)cuda_stream : h->stream, nullptr);
}

int tks_submit_host(tks_handle *h, const void *query, uint32_t k, uint32_t flags, uint64_t *ticket) {
    if (!h || !query || !ticket) return TKS_EINVAL;
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) {
        if (flags & TKS_SUBMIT_EXCHANGE) return h->fail(TKS_EINVAL, "BS-CSR mode spreads partitions, not candidates (ShardedSpMVFixed)");
        TKS_CUDA(h, cudaSetDevice(h->device));
        return bscsr_submit_host(h, static_cast<const uint32_t *>(query), k, ticket);
    }
    return pipe_submit(h, nullptr, static_cast<const float *>(query), k, flags, h->stream, ticket);
}

int tks_fetch(tks_handle *h, uint64_t ticket, uint32_t *idx_out, void *val_out_v, uint32_t *count) {
    if (!h || !idx_out || !val_out_v) return TKS_EINVAL;
    if (h->cfg.mode == TKS_MODE_FIXED_BSCSR) {
        TKS_CUDA(h, cudaSetDevice(h->device));
        return bscsr_fetch_ticket(h, ticket, idx_out, static_cast<uint32_t *>(val_out_v), count);
    }
    float *val_out = static_cast<float *>(val_out_v);
    if (!h->d_pipe_state || ticket == 0 || ticket > h->pipe_seq) return h->fail(TKS_EINVAL, "unknown ticket");
    const int slot = (int)(ticket % (uint64_t)h->pipe_slots);
    if (h->pipe_slot_ticket[slot] != ticket)
        return h->fail(TKS_ESTATE, "the result of ticket %llu is gone: at most %d queries are kept, fetch before submitting further",
                       (unsigned long long)ticket, h->pipe_slots);
    TKS_CUDA(h, cudaSetDevice(h->device));
    TKS_CUDA(h, cudaEventSynchronize(h->pipe_ev_done[slot]));
    h->pipe_busy[slot] = false;
    const uint32_t *blk = h->h_pipe_res + (size_t)slot * pipe_res_words(h);
    if (blk[0] == kPeerTimeout)
        return h->fail(TKS_ECUDA, "device-side wait timed out (TKS_SPIN_TIMEOUT_MS): a rank of the box never delivered its "
                                  "candidates, or a kernel of the pipelined submit never completed");
    const uint32_t k = h->pipe_slot_k[slot];
    std::memcpy(idx_out, blk + 64, k * 4u);
    std::memcpy(val_out, blk + 64 + h->kmax, k * 4u);
    if (count) *count = blk[0];
    return TKS_OK;
}

int tks_pipeline_wait(tks_handle *h, void *cuda_stream) {
    if (!h) return TKS_EINVAL;
    if (h->pipe_last_slot < 0) return TKS_OK;
    TKS_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    TKS_CUDA(h, cudaStreamWaitEvent(s, h->pipe_ev_done[h->pipe_last_slot], 0));
    return TKS_OK;
}

int tks_pipeline_stamps(tks_handle *h, uint64_t *stamps_ns, uint32_t capacity, uint32_t *count) {
    if (!h || !count) return TKS_EINVAL;
    *count = 0;
    if (!h->d_pipe_stamps) return TKS_OK;
    TKS_CUDA(h, cudaSetDevice(h->device));
    int rc = pipe_drain(h);
    if (rc) return rc;
    static_assert(kStampWords == TKS_PIPE_STAMP_WORDS, "header and kernels disagree on the stamp record");
    TKS_CUDA(h, cudaMemcpy(h->h_pipe_stamps, h->d_pipe_stamps, (size_t)tks::Handle::kPipeStamps * kStampWords * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    uint32_t n = h->pipe_seq < tks::Handle::kPipeStamps - 1 ? h->pipe_seq : tks::Handle::kPipeStamps - 1;
    if (n > capacity) n = capacity;
    for (uint32_t i = 0; i < n; i++) {   // oldest first
        const uint32_t seq = h->pipe_seq - n + 1u + i;
        if (stamps_ns)
            std::memcpy(stamps_ns + (size_t)i * kStampWords, h->h_pipe_stamps + (size_t)(seq % tks::Handle::kPipeStamps) * kStampWords,
                        kStampWords * sizeof(uint64_t));
    }
    *count = n;
    return TKS_OK;
}

int tks_set_profile_kernels(tks_handle *h, int on) {
    if (!h) return TKS_EINVAL;
    h->cfg.profile_kernels = on ? 1 : 0;
    return TKS_OK;
}

int tks_get_stats(tks_handle *h, tks_stats *out) {
    if (!h || !out) return TKS_EINVAL;
    if (h->cfg.mode == TKS_MODE_FLOAT_CSR && h->d_state) {
        cudaSetDevice(h->device);
        RunState st{};
        const RunState *src = (h->last_run_pipelined && h->pipe_last_slot >= 0) ? h->d_pipe_state + h->pipe_last_slot : h->d_state;
        if (cudaMemcpy(&st, src, sizeof st, cudaMemcpyDeviceToHost) == cudaSuccess)
            h->stats.last_candidates = st.result_count;
    }
    *out = h->stats;
    return TKS_OK;
}

}  // extern "C"

// ---- several GPUs driven by ONE process (SURVEY 8b `num_gpus / device_ids`, 8e) ------------------------------------
// A group owns one handle per device.  The devices map each other's exchange windows by plain peer access (one address
// space, no IPC), every shard runs the same three launches as a single device, and the select kernel of every shard
// exchanges its K candidates over NVLink and merges (csr_topk.cuh) -- afterwards every device holds the global top-k.

struct tks_group {
    std::vector<tks_handle *> h;
    std::vector<uint64_t> row_begin;   // first global row of every shard (+ total rows at the end)
    std::string err;
    uint32_t cols = 0;
    bool have_matrix = false;
    int fail(int code, const std::string &m) { err = m; return code; }
    int adopt(int rc, tks_handle *hh) { if (rc) err = hh->err; return rc; }
};

extern "C" {

const char *tks_group_last_error(const tks_group *g) { return g ? g->err.c_str() : g_create_error.c_str(); }

void tks_group_destroy(tks_group *g) {
    if (!g) return;
    for (tks_handle *hh : g->h) {
        if (!hh) continue;
        // the windows of the other members were never opened through IPC: forget them before the handle closes them
        for (uint32_t r = 0; r < kPeerMaxWorld; r++) hh->peer_mapped[r] = nullptr;
        hh->peer_world = 0;
        tks_destroy(hh);
    }
    delete g;
}

int tks_group_create(const tks_config *cfg, const int32_t *devices, uint32_t n, tks_group **out) {
    if (!cfg || !devices || !out) { g_create_error = "null argument"; return TKS_EINVAL; }
    *out = nullptr;
    if (n < 1 || n > kPeerMaxWorld) { g_create_error = "a group has 1..8 devices"; return TKS_EINVAL; }
    if (cfg->mode != TKS_MODE_FLOAT_CSR) { g_create_error = "groups run the float CSR engine (FPGA mode spreads its partitions with ShardedSpMVFixed)"; return TKS_EINVAL; }
    tks_group *g = new (std::nothrow) tks_group();
    if (!g) { g_create_error = "out of memory"; return TKS_ENOMEM; }
    g->h.assign(n, nullptr);
    for (uint32_t r = 0; r < n; r++) {
        tks_config c = *cfg;
        c.device = devices[r];
        int rc = tks_create(&c, &g->h[r]);
        if (rc) { tks_group_destroy(g); return rc; }   // g_create_error is set
    }
    // peer access between every pair of distinct devices, then the exchange windows
    for (uint32_t a = 0; a < n && n > 1; a++) {
        cudaSetDevice(devices[a]);
        for (uint32_t b = 0; b < n; b++) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (!can) { g_create_error = "devices of the group cannot access each other's memory"; tks_group_destroy(g); return TKS_ECUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { g_create_error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); tks_group_destroy(g); return TKS_ECUDA; }
        }
    }
    if (n > 1) {
        const size_t bytes = peer_window_bytes(kKMax);
        for (uint32_t r = 0; r < n; r++) {
            tks_handle *hh = g->h[r];
            cudaSetDevice(hh->device);
            if (cudaMalloc(&hh->d_peer_window, bytes) != cudaSuccess || cudaMemset(hh->d_peer_window, 0, bytes) != cudaSuccess) {
                g_create_error = "exchange window allocation failed"; tks_group_destroy(g); return TKS_ECUDA;
            }
        }
        for (uint32_t r = 0; r < n; r++) {
            tks_handle *hh = g->h[r];
            for (uint32_t o = 0; o < n; o++) hh->peer_mapped[o] = g->h[o]->d_peer_window;
            hh->peer_world = n; hh->peer_rank = r; hh->peer_seq = 0; hh->peer_ready = true;
        }
    }
    *out = g;
    return TKS_OK;
}

uint32_t tks_group_size(const tks_group *g) { return g ? (uint32_t)g->h.size() : 0u; }

tks_handle *tks_group_member(tks_group *g, uint32_t i) { return (g && i < g->h.size()) ? g->h[i] : nullptr; }

int tks_group_upload_csr(tks_group *g, uint64_t rows, uint32_t cols, uint64_t nnz, const void *ptr, int ptr_bits,
                         const uint32_t *idx, const float *val) {
    if (!g || !ptr || (nnz && (!idx || !val))) return TKS_EINVAL;
    if (ptr_bits != 32 && ptr_bits != 64) return g->fail(TKS_EINVAL, "ptr_bits must be 32 or 64");
    const uint32_t n = (uint32_t)g->h.size();
    auto at = [&](uint64_t r) -> uint64_t {
        return ptr_bits == 64 ? static_cast<const uint64_t *>(ptr)[r] : static_cast<const uint32_t *>(ptr)[r];
    };
    if (at(rows) != nnz) return g->fail(TKS_EINVAL, "ptr[rows] != nnz");
    // contiguous row shards balanced by non-zeros (the reference's partition rule, host_spmv_bscsr.cpp:136-141, with
    // boundaries balanced by nnz instead of by row count)
    g->row_begin.assign(n + 1, rows);
    g->row_begin[0] = 0;
    for (uint32_t s = 1; s < n; s++) {
        const uint64_t target = nnz / n * s;
        uint64_t lo = g->row_begin[s - 1], hi = rows;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (at(mid) < target) lo = mid + 1; else hi = mid; }
        g->row_begin[s] = lo;
    }
    for (uint32_t s = 0; s < n; s++) {
        const uint64_t r0 = g->row_begin[s], r1 = g->row_begin[s + 1], b = at(r0), e = at(r1);
        std::vector<uint64_t> p(r1 - r0 + 1);
        for (uint64_t r = r0; r <= r1; r++) p[r - r0] = at(r) - b;
        int rc = tks_upload_csr(g->h[s], r1 - r0, cols, e - b, p.data(), 64, idx + b, val + b, r0);
        if (rc) return g->adopt(rc, g->h[s]);
    }
    g->cols = cols;
    g->have_matrix = true;
    return TKS_OK;
}

int tks_group_generate_synthetic(tks_group *g, uint64_t rows, uint32_t cols, uint32_t avg_degree, int dist, uint64_t seed) {
    if (!g) return TKS_EINVAL;
    const uint32_t n = (uint32_t)g->h.size();
    g->row_begin.assign(n + 1, rows);
    for (uint32_t s = 0; s <= n; s++) g->row_begin[s] = rows / n * s + (s < rows % n ? s : rows % n);
    for (uint32_t s = 0; s < n; s++) {
        int rc = tks_generate_synthetic(g->h[s], g->row_begin[s + 1] - g->row_begin[s], cols, avg_degree, dist, seed, g->row_begin[s]);
        if (rc) return g->adopt(rc, g->h[s]);
    }
    g->cols = cols;
    g->have_matrix = true;
    return TKS_OK;
}

int tks_group_set_query(tks_group *g, const float *vec) {
    if (!g || !vec) return TKS_EINVAL;
    for (tks_handle *hh : g->h) {
        int rc = tks_set_query(hh, vec, 1);
        if (rc) return g->adopt(rc, hh);
    }
    return TKS_OK;
}

int tks_group_run(tks_group *g, uint32_t k, float *kernel_ms, float *total_ms) {
    if (!g) return TKS_EINVAL;
    if (!g->have_matrix) return g->fail(TKS_ESTATE, "no matrix uploaded");
    if (g->h.size() == 1) return g->adopt(tks_run(g->h[0], k, kernel_ms, total_ms), g->h[0]);
    auto t0 = std::chrono::high_resolution_clock::now();
    // every shard: sample -> main -> select (+ exchange over the peer windows + merge), enqueued back to back; the
    // devices run concurrently and meet inside their select kernels
    for (tks_handle *hh : g->h) {
        cudaSetDevice(hh->device);
        cudaEventRecord(hh->ev0, hh->stream);
        int rc = tks_run_exchange_async(hh, k, nullptr);
        if (rc) return g->adopt(rc, hh);
        cudaEventRecord(hh->ev1, hh->stream);
    }
    float worst = 0.f;
    for (tks_handle *hh : g->h) {
        cudaSetDevice(hh->device);
        cudaError_t e = cudaStreamSynchronize(hh->stream);
        if (e != cudaSuccess) return g->fail(TKS_ECUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, hh->ev0, hh->ev1);
        if (ms > worst) worst = ms;
    }
    const float tot = std::chrono::duration<float, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    for (tks_handle *hh : g->h) { hh->stats.last_kernel_ms = worst; hh->stats.last_total_ms = tot; }
    if (kernel_ms) *kernel_ms = worst;   // the slowest device, exchange and merge included
    if (total_ms) *total_ms = tot;
    return TKS_OK;
}

int tks_group_read_result(tks_group *g, uint32_t member, uint32_t *idx_out, float *val_out, uint32_t *count) {
    if (!g || member >= g->h.size()) return TKS_EINVAL;
    return g->adopt(tks_read_result(g->h[member], 0, idx_out, val_out, count), g->h[member]);
}

int tks_group_submit_host(tks_group *g, const float *query, uint32_t k, uint64_t *ticket) {
    if (!g || !query || !ticket) return TKS_EINVAL;
    if (!g->have_matrix) return g->fail(TKS_ESTATE, "no matrix uploaded");
    const uint32_t flags = g->h.size() > 1 ? TKS_SUBMIT_EXCHANGE : 0u;
    uint64_t t = 0;
    for (tks_handle *hh : g->h) {
        int rc = tks_submit_host(hh, query, k, flags, &t);
        if (rc) return g->adopt(rc, hh);
    }
    *ticket = t;   // the members are submitted to in lock step: one ticket names the step on all of them
    return TKS_OK;
}

int tks_group_fetch(tks_group *g, uint64_t ticket, uint32_t *idx_out, float *val_out, uint32_t *count) {
    if (!g) return TKS_EINVAL;
    // every member holds the global result; member 0's copy is returned, the others' slots are released
    for (size_t i = g->h.size(); i-- > 1;) {
        tks_handle *hh = g->h[i];
        const int slot = (int)(ticket % (uint64_t)hh->pipe_slots);
        cudaSetDevice(hh->device);
        if (hh->pipe_busy[slot] && hh->pipe_slot_ticket[slot] == ticket) {
            cudaEventSynchronize(hh->pipe_ev_done[slot]);
            hh->pipe_busy[slot] = false;
        }
    }
    return g->adopt(tks_fetch(g->h[0], ticket, idx_out, val_out, count), g->h[0]);
}

}  // extern "C"

