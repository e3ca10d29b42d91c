/*
 * topkspmv.h -- C ABI of libtopkspmv.so, the B200-native fused Top-K SpMV engine.
 *
 * The reference (AlbertoParravicini/approximate-spmv-topk) has no FFI layer; its
 * accelerator seam is the per-backend `struct SpMV` functor that every host main()
 * drives through four verbs (citations relative to the reference checkout):
 *
 *   construct      src/fpga/src/host_spmv_bscsr.cpp:104   src/gpu/host_spmv_topk_csr_gpu.cu:95
 *   operator()     src/fpga/src/host_spmv_bscsr.cpp:323   src/gpu/host_spmv_topk_csr_gpu.cu:171
 *   read_result    src/fpga/src/host_spmv_bscsr.cpp:399   src/gpu/host_spmv_topk_csr_gpu.cu:233
 *   reset          src/fpga/src/host_spmv_bscsr.cpp:450   src/gpu/host_spmv_topk_csr_gpu.cu:241
 *
 * Each entry point below names the reference interface it replaces.  Plain C,
 * pointers and sizes only; no C++/torch types cross this boundary.  All calls
 * return 0 on success, a negative TKS_E* code otherwise; tks_last_error() gives
 * the message.  One handle = one matrix shard resident on one CUDA device.
 * A handle is thread-compatible (one caller at a time), like the reference's SpMV.
 * There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef TOPKSPMV_H
#define TOPKSPMV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TKS_VERSION 1

/* error codes */
#define TKS_OK 0
#define TKS_EINVAL (-1)   /* bad argument / unsupported knob value          */
#define TKS_ECUDA (-2)    /* CUDA runtime error (message has the details)   */
#define TKS_ESTATE (-3)   /* call out of order (e.g. run before upload)     */
#define TKS_ENOMEM (-4)
#define TKS_EIO (-5)      /* file errors in the host-side loaders           */

/* mode: which reference back-end's semantics the handle reproduces */
#define TKS_MODE_FLOAT_CSR 0   /* exact fp32 CSR  (src/gpu/host_spmv_topk_csr_gpu.cu)                */
#define TKS_MODE_FIXED_BSCSR 1 /* FPGA semantics: W-bit fixed point, BS-CSR packets, P partitions x  */
                               /* LFR lanes x local K (src/fpga/src/ip/spmv/spmv_bscsr_top_k_*.{hpp,cpp}) */

/* storage type of the matrix values in float mode */
#define TKS_VALUE_FP32 0
#define TKS_VALUE_FP16 1
#define TKS_VALUE_BF16 2   /* bfloat16 storage; not a mode of the reference */

/* tie-break of equal scores in the final list */
#define TKS_TIE_LOWER_INDEX 0  /* north-star contract                                          */
#define TKS_TIE_HIGHER_INDEX 1 /* reference sort_tuples, src/common/utils/evaluation_utils.hpp:52-56 */

/* Runtime form of the reference's compile-time knobs (src/common/types.hpp). */
typedef struct tks_config {
    int32_t mode;                  /* TKS_MODE_*                                              */
    int32_t fixed_width;           /* FIXED_WIDTH  types.hpp:20   (17..32; BS-CSR mode only)  */
    int32_t partitions;            /* SPMV_PARTITIONS types.hpp:36 (BS-CSR mode only)         */
    int32_t local_k;               /* K  types.hpp:51  per-lane local top-K (BS-CSR mode)     */
    int32_t limited_finished_rows; /* LIMITED_FINISHED_ROWS types.hpp:77 (BS-CSR mode)        */
    int32_t max_cols;              /* MAX_COLS types.hpp:55 (1024; float mode accepts <=16383) */
    int32_t tie_break;             /* TKS_TIE_*                                               */
    int32_t device;                /* CUDA device ordinal                                     */
    int32_t max_batch;             /* max queries per run (float mode), >= 1                  */
    int32_t chunk_nnz;             /* float mode work-unit size in nnz (0 = default)          */
    int32_t profile_kernels;       /* 1: tks_run also times the dominant kernel alone (stats)  */
    int32_t batch_mode;            /* float mode, batch > 1: 0 = one matrix pass per 32 queries     */
                                   /* (csr_batched.cuh; needs cols <= ~1500), 1 = one pass per query */
    int32_t batch_pool_cap;        /* candidate keys per query in batched mode (0 = 32768)     */
    int32_t batch_fma;             /* batched mode: 1 = fused multiply-add (scores no longer   */
                                   /* bit-identical to the sequential fp32 gold)               */
    int32_t fixed_drift_free;      /* BS-CSR mode, FIXED_WIDTH <= 22 and LFR >= 2: 1 = repair the reference's    */
                                   /* row-counter drift: packets with more than LFR row segments   */
                                   /* advance the row counter by the true row count and carry the true last      */
                                   /* partial sum; 0 = the reference's semantics, bit for bit (default)          */
    int32_t value_type;            /* float mode: TKS_VALUE_FP32 (default) or TKS_VALUE_FP16 = the reference's          */
                                   /* half-precision GPU mode (-a, options.hpp:82; host_spmv_topk_csr_gpu.cu:132-136,   */
                                   /* 151-153): matrix values and query rounded to IEEE half, fp32 accumulation;        */
                                   /* TKS_VALUE_BF16: the same with bfloat16 storage                                    */
} tks_config;

typedef struct tks_handle tks_handle;

/* What the last run moved and how long it took (algorithmic bytes). */
typedef struct tks_stats {
    uint64_t rows, cols, nnz;
    uint64_t packets;             /* BS-CSR mode: total 64-byte packets                 */
    uint64_t algorithmic_bytes;   /* per query, formula of DESIGN.md "Roofline"         */
    uint64_t device_bytes;        /* bytes actually resident for the matrix              */
    float last_kernel_ms;         /* device time of the last tks_run (CUDA events)       */
    float last_total_ms;          /* host wall time of the last tks_run                  */
    uint32_t last_candidates;     /* candidates that survived the threshold filter       */
    uint32_t launches_per_run;    /* kernels launched by one tks_run                     */
    float last_main_kernel_ms;    /* dominant kernel alone (only with cfg.profile_kernels) */
    uint32_t batched_fallbacks;   /* batched queries re-run alone because their pool overflowed */
    uint32_t logged_candidates;   /* BS-CSR mode with profile_kernels: entries the stream kernel logged */
    uint32_t work_unit_nnz;       /* float mode: non-zeros per work unit of the resident matrix */
    uint32_t work_units;          /* float mode: number of work units                                       */
    uint32_t reserved[3];
} tks_stats;

/* ---- lifecycle ---------------------------------------------------------- */

int tks_default_config(tks_config *cfg);                       /* types.hpp defaults          */
int tks_create(const tks_config *cfg, tks_handle **out);       /* SpMV ctor, first half       */
void tks_destroy(tks_handle *h);
const char *tks_last_error(const tks_handle *h);               /* h may be NULL (create errors) */
int tks_version(void);

/* ---- matrix upload (SpMV ctor, second half: host:133,250-321; gpu:105-168) */

/* CSR as built by coo2csr (src/common/utils/utils.hpp:522-580).  ptr has rows+1
 * entries of 32 or 64 bits; idx/val have nnz entries.  row_offset is added to
 * every reported index (global row id of local row 0, for row-sharded use).  */
int tks_upload_csr(tks_handle *h, uint64_t rows, uint32_t cols, uint64_t nnz, const void *ptr,
                   int ptr_bits, const uint32_t *idx, const float *val, uint64_t row_offset);
/* Same, with DEVICE pointers (data already in HBM, e.g. generated there). */
int tks_upload_csr_device(tks_handle *h, uint64_t rows, uint32_t cols, uint64_t nnz,
                          const void *d_ptr, int ptr_bits, const uint32_t *d_idx,
                          const float *d_val, uint64_t row_offset);

/* Caller-built BS-CSR packets (host_spmv_bscsr.cpp:189-248 output): for each of
 * `partitions` partitions a host array of packets_per_part[p] 64-byte words,
 * the partition's first row (host:145) and its number of COO entries.         */
int tks_upload_bscsr(tks_handle *h, uint32_t cols, uint32_t partitions,
                     const uint64_t *packets_per_part, const void *const *packets,
                     const uint32_t *first_row, const uint64_t *nnz_per_part);

/* GPU-side packer: what the reference FPGA constructor does on the host with one thread --
 * row partitioning (host_spmv_bscsr.cpp:112-121,136-150), packet_coo / packet_coo_partition (:133-248) with
 * the bit layout of fpga_utils.hpp:307-365 -- plus this engine's device tables, all on the device.
 * row/col/val32: row-sorted COO, val32 = raw ap_ufixed<32,1> words (the reference ctor's x, y, val arguments,
 * host:104).  Leaves exactly the resident state of tks_pack_bscsr + tks_upload_bscsr.  BS-CSR mode only.   */
int tks_upload_coo_fixed(tks_handle *h, const uint32_t *row, const uint32_t *col, const uint32_t *val32,
                         uint64_t nnz, uint32_t num_rows, uint32_t cols);
/* Same, with DEVICE pointers. */
int tks_upload_coo_fixed_device(tks_handle *h, const uint32_t *d_row, const uint32_t *d_col,
                                const uint32_t *d_val32, uint64_t nnz, uint32_t num_rows, uint32_t cols);
/* FNV-1a digests (16 words) of the resident BS-CSR state -- packets, chunk tables, sample tables, first rows --
 * so that a checker can assert two upload paths are byte-identical.                                      */
int tks_bscsr_state_digest(tks_handle *h, uint64_t *digest, uint32_t n);

/* Synthetic matrix generated in HBM with the law of
 * src/resources/python/create_matrices.py:84-104 (dist 0 = uniform, 1 = gamma).
 * Float mode only.  row_offset/global_rows let N ranks generate disjoint shards. */
int tks_generate_synthetic(tks_handle *h, uint64_t rows, uint32_t cols, uint32_t avg_degree,
                           int dist, uint64_t seed, uint64_t row_offset);

/* Copy the resident CSR back to the host (ptr64: rows+1, idx/val: nnz; any may be NULL).
 * Lets a checker see exactly the matrix tks_generate_synthetic made.           */
int tks_download_csr(tks_handle *h, uint64_t *ptr64, uint32_t *idx, float *val);
/* The same for the rows [row_begin, row_end) only: ptr64 gets row_end - row_begin + 1 offsets into the WHOLE matrix
 * (so ptr64[last] - ptr64[0] non-zeros follow in idx / val).  Call once with idx = val = NULL to size the arrays.
 * Lets a checker look at a bounded sample of a shard too large to copy back (BASELINE config 4).               */
int tks_download_csr_rows(tks_handle *h, uint64_t row_begin, uint64_t row_end, uint64_t *ptr64, uint32_t *idx,
                          float *val);

/* ---- per query ---------------------------------------------------------- */

/* reset(vec): float mode: batch x cols fp32 values, row-major.  BS-CSR mode:
 * cols raw 32-bit ap_ufixed<32,1> words (what write_block_vec stores,
 * src/fpga/src/ip/fpga_utils.hpp:346-355), batch must be 1.                   */
int tks_set_query(tks_handle *h, const void *vec, uint32_t batch);
int tks_set_query_device(tks_handle *h, const void *d_vec, uint32_t batch, void *cuda_stream);

/* operator(): launch and wait.  k = CLI -k (options.hpp:103).  Timings optional. */
int tks_run(tks_handle *h, uint32_t k, float *kernel_ms, float *total_ms);
/* Enqueue only (no sync) on the caller's stream; result stays on the device.
 * Every `cuda_stream` argument of this header is a cudaStream_t; NULL means the
 * handle's PRIVATE (non-blocking) stream, not the default stream -- a caller
 * whose work sits on the default stream names it with cudaStreamLegacy or
 * cudaStreamPerThread (the Python layer does this for torch's default stream). */
int tks_run_async(tks_handle *h, uint32_t k, void *cuda_stream);

/* read_result(): sorted (score desc, tie-break).  Float mode: val_out is
 * float[k].  BS-CSR mode: val_out is uint32_t[] raw ap_ufixed<32,1> scores and
 * *count may be < k (host_spmv_bscsr.cpp:399-448).  Buffers must hold k entries. */
int tks_read_result(tks_handle *h, uint32_t query, uint32_t *idx_out, void *val_out, uint32_t *count);

/* BS-CSR mode: the kernel's raw output in the reference layout
 * (spmv_bscsr_top_k_multicore.cpp:151-185): partitions x local_k words of
 * 16 x u32 (row index local to the partition) and 16 x u32 (ap_ufixed<32,1>). */
int tks_read_partition_results(tks_handle *h, uint32_t *idx_words, uint32_t *val_words);

/* BS-CSR mode, several GPUs: device address of the result words of the last tks_run_async (index words, then value
 * words, one block of *n_words 32-bit words), so that ranks can all-gather them device to device before the host
 * merge (tks_merge_partition_words).                                                                            */
int tks_partition_words_device(tks_handle *h, const uint32_t **d_words, uint32_t *n_words);

/* ---- multi-GPU plumbing: K candidates per rank, merged after an all-gather */

/* Device pointer to the last run's sorted candidates of query q as 64-bit keys
 * (score bits << 32 | tie-ordered index) and their count (<= k).               */
int tks_result_keys_device(tks_handle *h, uint32_t query, const uint64_t **d_keys, uint32_t *count);
/* Merge n_keys gathered keys (device) into this handle's result for query q.   */
int tks_merge_keys_device(tks_handle *h, uint32_t query, const uint64_t *d_keys, uint32_t n_keys,
                          uint32_t k, void *cuda_stream);

/* Batched form: d_keys is [batch][keys_per_query] (the all-gathered candidates regrouped by query);
 * one launch merges every query.                                               */
int tks_merge_keys_batched_device(tks_handle *h, const uint64_t *d_keys, uint32_t keys_per_query,
                                  uint32_t batch, uint32_t k, void *cuda_stream);

/* Candidate exchange over PEER MEMORY (NVLink / NVSwitch), one process per GPU: every rank maps every other rank's
 * exchange window through CUDA IPC; tks_run_exchange_async then runs the query on the local shard with the SAME three
 * launches as tks_run_async, but the select kernel stores its k candidates into every peer's window, waits for
 * theirs and merges -- no NCCL launch and no merge launch per query; afterwards every rank holds the global top-k.
 * Protocol: tks_peer_init on every rank (fills a TKS_IPC_HANDLE_BYTES blob), exchange the blobs by any means
 * (torch.distributed all_gather_object), tks_peer_connect with all of them in rank order.  Float mode, one query
 * per run, world <= 8, world * k <= 2048; every rank must make the same call for the same step.  A peer that never
 * delivers makes tks_read_result fail after a bounded wait instead of hanging the device.                         */
#define TKS_IPC_HANDLE_BYTES 128
int tks_peer_init(tks_handle *h, uint32_t world, uint32_t rank, void *ipc_handle_out);
int tks_peer_connect(tks_handle *h, const void *all_handles);
int tks_run_exchange_async(tks_handle *h, uint32_t k, void *cuda_stream);
int tks_peer_exchange_async(tks_handle *h, uint32_t k, void *cuda_stream);  /* exchange + merge as a stand-alone
                                                                              launch after tks_run_async     */

/* ---- pipelined submits: consecutive queries overlap (float mode, one query per submit) -------------------------
 * The reference hosts loop  reset(vec) -> operator() -> read_result()  strictly one query after the other
 * (src/gpu/host_spmv_topk_csr_gpu.cu:399-423); a caller that has the next query ready before it needs the previous
 * result can instead SUBMIT queries: tks_submit enqueues the three kernels of one query on three streams -- the
 * threshold sample on an engine stream (it runs beside the main kernel of the previous query), the matrix stream on
 * `cuda_stream` (chained to the previous query's by programmatic dependent launch without waiting for it, so the
 * stream of non-zeros never pauses between queries), the select (+ peer exchange + merge with TKS_SUBMIT_EXCHANGE,
 * see tks_peer_init) on a second engine stream -- with per-slot scratch and sequence-number hand-overs in HBM instead
 * of stream order.  Results are those of tks_run_async for the same query, bit for bit.
 *   d_query  DEVICE pointer to cols fp32 values; it is read in place and must stay valid and unchanged until the
 *            query's result is complete (tks_pipeline_wait / tks_read_result).
 *   flags    TKS_SUBMIT_QUERY_READY: the query's bytes are already complete in memory at the time of the call
 *            (otherwise the sample stream is made to wait for the work enqueued on `cuda_stream` so far);
 *            TKS_SUBMIT_EXCHANGE: several GPUs -- every rank submits the same step (tks_run_exchange_async rules).
 * At most four queries are in flight: a fifth submit blocks the host until the first one's select has finished.
 * tks_read_result (synchronises) returns the LAST submitted query's result; tks_pipeline_wait makes `cuda_stream`
 * wait for it on the device without blocking the host.  tks_pipeline_stamps returns, for each of the last `capacity`
 * submits (oldest first), TKS_PIPE_STAMP_WORDS %globaltimer nanosecond stamps written by the kernels themselves:
 * [0] sample begin, [1] sample end, [2] main kernel begins streaming, [3] main kernel's last CTA done, [4] select CTA
 * resident, [5] select begins (pool complete), [6] select (+ exchange + merge) done -- the pipeline's timeline
 * without any event in the streams.  stamps_ns holds capacity * TKS_PIPE_STAMP_WORDS words.                   */
#define TKS_PIPE_STAMP_WORDS 8
/* The same pipeline fed from and read back into HOST memory -- the throughput form of reset(vec) + operator() +
 * read_result(): tks_submit_host copies `query` (cols fp32 values, any host memory; it may be reused as soon as the
 * call returns) to the device on the sample stream and returns a ticket; the select kernel of that query stores the
 * sorted indices, scores and the count straight into a pinned host block; tks_fetch(ticket) waits for that query alone
 * and copies its first k results out (val_out float[k]).  Results are kept for the last four tickets: fetch ticket t
 * before submitting ticket t + 4.  Host-to-device and device-to-host traffic of every query is part of its step.
 * BS-CSR mode: `query` is the 1024 raw ap_ufixed<32,1> words of tks_set_query, val_out receives raw uint32 scores and
 * *count may be < k (read_result semantics of host_spmv_bscsr.cpp:399-448); two queries are kept; the sample of query
 * i + 1 (and its copy) overlaps the stream and replay kernels of query i, and the host merge of query i runs in
 * tks_fetch while the device works on query i + 1.                                                              */
int tks_submit_host(tks_handle *h, const void *query, uint32_t k, uint32_t flags, uint64_t *ticket);
int tks_fetch(tks_handle *h, uint64_t ticket, uint32_t *idx_out, void *val_out, uint32_t *count);

#define TKS_SUBMIT_EXCHANGE 1u
#define TKS_SUBMIT_QUERY_READY 2u
int tks_submit(tks_handle *h, const void *d_query, uint32_t k, uint32_t flags, void *cuda_stream);   /* BS-CSR mode: 1024 raw words, sample overlap only */
int tks_pipeline_wait(tks_handle *h, void *cuda_stream);
int tks_pipeline_stamps(tks_handle *h, uint64_t *stamps_ns, uint32_t capacity, uint32_t *count);

/* ---- several GPUs driven by ONE process (SURVEY 8b num_gpus / device_ids, 8e) ------------------------------------
 * The counterpart of the one-process-per-GPU plumbing above for a plain C/C++ host such as the reference's main()
 * (src/gpu/host_spmv_topk_csr_gpu.cu:291-480): a group owns one engine per device, the devices map each other's
 * exchange windows by peer access (no IPC, no NCCL, no torch), rows are sharded contiguously (balanced by non-zeros),
 * every shard runs the usual three launches and the select kernels exchange the K candidates over NVLink and merge.
 * After a run every member holds the GLOBAL top-k.  The same device may be listed more than once (two shards on one
 * GPU): that exercises the whole exchange on a single-GPU box.  Float CSR mode, world * k <= 2048.               */
typedef struct tks_group tks_group;
int tks_group_create(const tks_config *cfg, const int32_t *devices, uint32_t n, tks_group **out);   /* cfg->device is ignored */
void tks_group_destroy(tks_group *g);
const char *tks_group_last_error(const tks_group *g);
uint32_t tks_group_size(const tks_group *g);
tks_handle *tks_group_member(tks_group *g, uint32_t i);                         /* e.g. for tks_get_stats */
int tks_group_upload_csr(tks_group *g, uint64_t rows, uint32_t cols, uint64_t nnz, const void *ptr, int ptr_bits,
                         const uint32_t *idx, const float *val);
int tks_group_generate_synthetic(tks_group *g, uint64_t rows, uint32_t cols, uint32_t avg_degree, int dist, uint64_t seed);
int tks_group_set_query(tks_group *g, const float *vec);                        /* reset(vec) on every member      */
int tks_group_run(tks_group *g, uint32_t k, float *kernel_ms, float *total_ms); /* operator(): launch all, wait    */
int tks_group_read_result(tks_group *g, uint32_t member, uint32_t *idx_out, float *val_out, uint32_t *count);
int tks_group_submit_host(tks_group *g, const float *query, uint32_t k, uint64_t *ticket);   /* pipelined, see above */
int tks_group_fetch(tks_group *g, uint64_t ticket, uint32_t *idx_out, float *val_out, uint32_t *count);

/* Switch tks_config.profile_kernels at run time: while on, tks_run brackets the dominant kernel with two extra
 * events (tks_stats.last_main_kernel_ms) and launches the kernels without overlap.                              */
int tks_set_profile_kernels(tks_handle *h, int on);

int tks_get_stats(tks_handle *h, tks_stats *out);

/* ---- host-side surface the reference hosts use around the accelerator ---- */

/* BSCSR_PACKET_SIZE (types.hpp:71-72) for a given FIXED_WIDTH.                 */
int tks_bscsr_packet_size(int fixed_width);
/* Value quantisation chain double -> ap_ufixed<32,1> -> float -> ap_ufixed<W,1>
 * (utils.hpp:401; fpga_utils.hpp:336-338).                                     */
uint32_t tks_fixed32_from_double(double v);
uint32_t tks_fixedW_from_fixed32(uint32_t raw32, int fixed_width);

/* Row partitioning + packet builder (host_spmv_bscsr.cpp:133-248) for row-sorted
 * COO with raw ap_ufixed<32,1> values.  Call once with packets == NULL to get
 * packets_per_part/first_row/nnz_per_part, then with `packets` pointing at
 * sum(packets_per_part) * 64 bytes, partitions laid out back to back.          */
int tks_pack_bscsr(const uint32_t *row, const uint32_t *col, const uint32_t *val32, uint64_t nnz,
                   uint32_t num_rows, int partitions, int fixed_width, uint64_t *packets_per_part,
                   uint32_t *first_row, uint64_t *nnz_per_part, void *packets);

/* read_result's merge (host_spmv_bscsr.cpp:399-448 + sort_tuples, evaluation_utils.hpp:40-62) over caller-held result
 * words in the kernel's layout (partitions x local_k words of 16 x u32, as tks_read_partition_results returns them):
 * idx += first_row[p], val > 0 kept, first insertion of an index wins, sorted (val desc, tie_break).  Lets a caller
 * merge the partitions of SEVERAL devices (FPGA mode, partitions spread over GPUs) exactly like one.
 * *count = number of distinct candidates; idx_out/val_out (k entries, may be NULL) get the first min(k, count).  */
int tks_merge_partition_words(uint32_t partitions, uint32_t local_k, uint32_t packet_size, const uint32_t *idx_words,
                              const uint32_t *val_words, const uint32_t *first_row, int tie_break, uint32_t k,
                              uint32_t *idx_out, uint32_t *val_out, uint32_t *count);

/* readMtx (utils.hpp:474-520 + mmio.hpp): coordinate real/integer/pattern general.
 * zero_indexed: the file's indices are 0-based (the reference hosts hard-code
 * true, host_spmv_bscsr.cpp:539; the MTX standard and its generator are 1-based).
 * Two-call protocol: nnz_capacity == 0 only fills rows/cols/nnz.              */
int tks_read_mtx(const char *path, int zero_indexed, int sort_tuples, int ignore_values,
                 uint32_t *rows, uint32_t *cols, uint64_t *nnz, uint64_t nnz_capacity,
                 uint32_t *x, uint32_t *y, double *val);

/* coo2csr (utils.hpp:522-580), stable counting sort by row.                    */
int tks_coo2csr(const uint32_t *x, const uint32_t *y, const float *val, uint64_t nnz, uint32_t rows,
                uint32_t cols, uint32_t *ptr, uint32_t *idx, float *out_val);

/* ---- binary matrix cache: skip re-parsing the MTX text (readMtx, utils.hpp:474-520) and
 * re-packing the BS-CSR packets (host_spmv_bscsr.cpp:133-248) on every run of a sweep.  Versioned,
 * checksummed container; truncated or altered files fail with TKS_EIO.  Read calls follow the two-call
 * protocol: all output arrays NULL only fills the sizes.                                           */
int tks_cache_write_csr(const char *path, uint64_t rows, uint32_t cols, uint64_t nnz, const uint64_t *ptr64,
                        const uint32_t *idx, const float *val);
int tks_cache_read_csr(const char *path, uint64_t *rows, uint32_t *cols, uint64_t *nnz, uint64_t *ptr64,
                       uint32_t *idx, float *val);
/* The same with a tag that says what the cache was made from: tks_cache_source_tag hashes the path, size and
 * modification time of the Matrix-Market file and the loader flags that change what the text parses to (index base,
 * ignore values).  A reader compares the stored tag with the tag of the matrix it was asked for and ignores a cache
 * that is stale or belongs to another file (0 = the cache is untagged).                                         */
uint64_t tks_cache_source_tag(const char *source_path, int zero_indexed, int ignore_values);
int tks_cache_write_csr_tagged(const char *path, uint64_t rows, uint32_t cols, uint64_t nnz, const uint64_t *ptr64,
                               const uint32_t *idx, const float *val, uint64_t source_tag);
int tks_cache_read_csr_tagged(const char *path, uint64_t *rows, uint32_t *cols, uint64_t *nnz, uint64_t *ptr64,
                              uint32_t *idx, float *val, uint64_t *source_tag);
int tks_cache_write_bscsr(const char *path, uint32_t rows, uint32_t cols, int fixed_width, uint32_t partitions,
                          const uint64_t *packets_per_part, const uint32_t *first_row,
                          const uint64_t *nnz_per_part, const void *packets);
int tks_cache_read_bscsr(const char *path, uint32_t *rows, uint32_t *cols, int *fixed_width, uint32_t *partitions,
                         uint64_t *total_packets, uint64_t *packets_per_part, uint32_t *first_row,
                         uint64_t *nnz_per_part, void *packets);

#ifdef __cplusplus
}
#endif
#endif /* TOPKSPMV_H */
