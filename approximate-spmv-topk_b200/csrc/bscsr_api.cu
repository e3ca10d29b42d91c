// bscsr_api.cu -- host side of the FPGA-semantics engine: packet upload + chunk tables, query
// transform, kernel dispatch on (FIXED_WIDTH, LIMITED_FINISHED_ROWS), result words, host merge.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "bscsr_pack.cuh"
#include "bscsr_topk.cuh"
#include "handle.hpp"

namespace tks {

struct BscsrState {
    uint32_t P = 0, cols = 0, B = 0;
    uint64_t total_packets = 0, total_nnz = 0;
    std::vector<uint32_t> first_row;
    uint8_t *d_packets = nullptr;
    uint32_t *d_chunk_first = nullptr, *d_chunk_count = nullptr, *d_chunk_local0 = nullptr, *d_chunk_row_in = nullptr,
             *d_chunk_lookback = nullptr, *d_chunk_part = nullptr, *d_part_chunk_begin = nullptr;
    uint32_t n_chunks = 0, chunk_cap = 0, tail_div = 4;
    BscsrLogs logs{};
    // sample pieces (first kBsSamplePackets packets of every partition, kBsSamplePiece each)
    uint32_t *d_s_first = nullptr, *d_s_count = nullptr, *d_s_local0 = nullptr, *d_s_lookback = nullptr, *d_s_part = nullptr,
             *d_s_part_begin = nullptr, *d_piece_top = nullptr, *d_ticket = nullptr, *d_theta_seed = nullptr,
             *d_sample_end = nullptr;   // [P] packets of the partition's prefix the sample covers
    uint32_t n_pieces = 0;
    uint32_t *d_xq = nullptr;        // 1024 pre-shifted query words
    uint32_t *h_xq = nullptr;        // pinned
    uint32_t *d_counter = nullptr;   // dynamic chunk scheduler
    uint32_t *d_res_idx = nullptr, *d_res_val = nullptr;   // P x Kp x 16 words each
    uint32_t *h_res_idx = nullptr, *h_res_val = nullptr;   // pinned
    bool have_query = false, have_words = false;
    int grid = 0;
    int variant = 0;             // TKS_BSCSR_VARIANT (experiments): 0 = default, 1 / 8 / 16 = query copies
    bool variant_ready = false;  // launch geometry computed
    bool replay_ready = false;
    cudaEvent_t ev_query = nullptr;
    bool bsx = false;            // packets re-encoded into the BSX device format (FIXED_WIDTH <= 22)
    // pipelined submits (tks_submit_host / tks_fetch in BS-CSR mode): the query of step i+1 is transformed, copied and
    // SAMPLED on a second stream while the stream and replay kernels of step i run, and the replay kernel writes the
    // result words straight into a pinned host block, so a step costs stream + replay only (bscsr_submit_host)
    static constexpr int kSlots = 2;
    cudaStream_t p_sample_stream = nullptr;
    cudaEvent_t p_ev_sample[kSlots] = {}, p_ev_done[kSlots] = {};
    uint32_t *p_d_xq[kSlots] = {}, *p_h_xq[kSlots] = {}, *p_d_theta[kSlots] = {}, *p_h_words[kSlots] = {};
    bool p_busy[kSlots] = {};
    uint64_t p_ticket[kSlots] = {};
    uint32_t p_k[kSlots] = {};
    uint64_t p_seq = 0;
    std::vector<uint32_t> merged_idx, merged_val;          // read_result() output of the last run
};

namespace {

template <int W, int LFR, int XREP, int THREADS, bool PREFETCH, bool BSX>
cudaError_t prep_stream_variant(int *ctas_per_sm) {
    auto kern = bscsr_stream_kernel<W, LFR, XREP, THREADS, PREFETCH, BSX>;
    const size_t smem = bscsr_stream_smem(XREP, THREADS);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (const char *v = std::getenv("TKS_BSCSR_CARVEOUT")) {   // experiment: L1 / shared split of the stream kernel, percent
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(v));
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kern, THREADS, smem);
}

// Where one run reads its query and seeds and writes its result words.  sample_stream != nullptr: the sample kernel
// runs there (the query and the seeds of the slot belong to it) and the main stream joins it through ev_sample.
struct BsRun {
    const uint32_t *xq;
    uint32_t *theta_seed;
    uint32_t *res_idx, *res_val;
    cudaStream_t sample_stream;
    cudaEvent_t ev_sample;
};

template <int W, int LFR, int XREP, int THREADS, bool PREFETCH, bool BSX>
void launch_stream_variant(Handle *h, BscsrState *b, const BscsrChunks &m, cudaStream_t s, const BsRun &r) {
    if (!b->variant_ready) {
        int per_sm = 1;
        cudaError_t e = prep_stream_variant<W, LFR, XREP, THREADS, PREFETCH, BSX>(&per_sm);
        if (e != cudaSuccess || per_sm < 1) per_sm = 1;
        b->grid = h->num_sms * per_sm;
        const uint32_t wpc = THREADS / 32;
        if ((uint32_t)b->grid * wpc > b->n_chunks) b->grid = (int)((b->n_chunks + wpc - 1) / wpc);
        b->variant_ready = true;
    }
    // programmatic dependent launch: the grid starts (and stages its query copies) while the sample kernel still runs
    const bool pdl = pdl_enabled() && !(h->cfg.profile_kernels != 0 && s == h->stream);
    launch_pdl(bscsr_stream_kernel<W, LFR, XREP, THREADS, PREFETCH, BSX>, dim3(b->grid), dim3(THREADS),
               bscsr_stream_smem(XREP, THREADS), s, pdl && r.sample_stream == nullptr, (const uint8_t *)b->d_packets, m, r.xq,
               (uint32_t)h->cfg.local_k, b->logs, (const uint32_t *)r.theta_seed, (const uint32_t *)b->d_sample_end,
               b->d_counter);
}

template <int W, int LFR, bool BSX>
void launch_stream_fmt(Handle *h, BscsrState *b, const BscsrChunks &m, cudaStream_t s, const BsRun &r) {
    const bool prof = h->cfg.profile_kernels != 0 && s == h->stream && r.sample_stream == nullptr;
    BscsrSample sm{b->d_s_first, b->d_s_count, b->d_s_local0, b->d_s_lookback, b->d_s_part, b->n_pieces,
                   b->d_s_part_begin, b->d_piece_top, b->d_ticket, r.theta_seed};
    const uint32_t sgrid = (b->n_pieces * 32u + kBsThreads - 1) / kBsThreads;
    if (r.sample_stream) {
        // pipelined submit: the sample of this query ran (or runs) on the sample stream, beside the previous query's kernels
        bscsr_sample_kernel<W, LFR, BSX><<<sgrid, kBsThreads, 0, r.sample_stream>>>(b->d_packets, sm, r.xq, (uint32_t)h->cfg.local_k);
        cudaEventRecord(r.ev_sample, r.sample_stream);
        cudaStreamWaitEvent(s, r.ev_sample, 0);
    } else {
        bscsr_sample_kernel<W, LFR, BSX><<<sgrid, kBsThreads, 0, s>>>(b->d_packets, sm, r.xq, (uint32_t)h->cfg.local_k);
    }
    if (prof) cudaEventRecord(h->evm0, s);
    // stream-kernel variants (query copies x CTA size); the alternatives exist for the headline format only
    if (W == 20 && LFR == 4 && BSX && b->variant == 1) launch_stream_variant<20, 4, 1, 256, false, BSX && W == 20>(h, b, m, s, r);
    else if (W == 20 && LFR == 4 && BSX && b->variant == 16) launch_stream_variant<20, 4, 16, 1024, false, BSX && W == 20>(h, b, m, s, r);
    else if (W == 20 && LFR == 4 && BSX && b->variant == 33) launch_stream_variant<20, 4, 32, 768, false, BSX && W == 20>(h, b, m, s, r);
    else if (W == 20 && LFR == 4 && BSX && b->variant == 34) launch_stream_variant<20, 4, 32, 1024, false, BSX && W == 20>(h, b, m, s, r);
    else if (W == 20 && LFR == 4 && BSX && b->variant == 35) launch_stream_variant<20, 4, 32, 1024, true, BSX && W == 20>(h, b, m, s, r);
    else if (W == 20 && LFR == 4 && BSX && b->variant == 36) launch_stream_variant<20, 4, 32, 896, true, BSX && W == 20>(h, b, m, s, r);
    else launch_stream_variant<W, LFR, kBsDefaultXrep, kBsDefaultThreads, true, BSX>(h, b, m, s, r);
    if (prof) cudaEventRecord(h->evm1, s);
}

template <int W, int LFR>
void launch_stream(Handle *h, BscsrState *b, const BscsrChunks &m, cudaStream_t s, const BsRun &r) {
    if constexpr (W + 10 <= 32) {
        if (b->bsx) { launch_stream_fmt<W, LFR, true>(h, b, m, s, r); return; }
    }
    launch_stream_fmt<W, LFR, false>(h, b, m, s, r);
}

template <int W>
int dispatch_lfr(Handle *h, BscsrState *b, const BscsrChunks &m, cudaStream_t s, const BsRun &r) {
    switch (h->cfg.limited_finished_rows) {
        case 1: launch_stream<W, 1>(h, b, m, s, r); break;
        case 2: launch_stream<W, 2>(h, b, m, s, r); break;
        case 3: launch_stream<W, 3>(h, b, m, s, r); break;
        case 4: launch_stream<W, 4>(h, b, m, s, r); break;
        default: return h->fail(TKS_EINVAL, "limited_finished_rows=%d is not instantiated (1..4)", h->cfg.limited_finished_rows);
    }
    if (!b->replay_ready) {
        cudaFuncSetAttribute(bscsr_replay_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kReplayDynSmem);
        b->replay_ready = true;
    }
    launch_pdl(bscsr_replay_kernel<W>, dim3(b->P * (uint32_t)h->cfg.limited_finished_rows), dim3(kReplayThreads),
               (size_t)kReplayDynSmem, s, pdl_enabled() && !(h->cfg.profile_kernels != 0 && s == h->stream), b->logs,
               (const uint32_t *)b->d_part_chunk_begin, (uint32_t)h->cfg.limited_finished_rows, (uint32_t)h->cfg.local_k,
               b->chunk_cap, r.res_idx, r.res_val, b->d_counter);
    return TKS_OK;
}

bool width_supported(int W) { return W == 20 || W == 21 || W == 25 || W == 26 || W == 32; }

}  // namespace

// Common first half of every upload path: knob validation and a fresh state object.
static int bscsr_begin_upload(Handle *h, uint32_t cols, uint32_t partitions, BscsrState **out) {
    if (h->cfg.mode != TKS_MODE_FIXED_BSCSR) return h->fail(TKS_ESTATE, "handle is not in FIXED_BSCSR mode");
    if (partitions != (uint32_t)h->cfg.partitions) return h->fail(TKS_EINVAL, "partitions != cfg.partitions");
    if (cols == 0 || cols > 1024) return h->fail(TKS_EINVAL, "cols outside 1..1024 (10-bit column field)");
    const int W = h->cfg.fixed_width, LFR = h->cfg.limited_finished_rows, Kp = h->cfg.local_k;
    if (!width_supported(W)) return h->fail(TKS_EINVAL, "fixed_width=%d is not instantiated (20, 21, 25, 26, 32)", W);
    if (Kp < 1 || Kp > (int)kBsMaxKp) return h->fail(TKS_EINVAL, "local_k outside 1..32");
    if (LFR < 1 || LFR > (int)kBsMaxLfr) return h->fail(TKS_EINVAL, "limited_finished_rows outside 1..4");
    bscsr_destroy(h);
    BscsrState *b = new BscsrState();
    h->bs = b;
    b->P = partitions; b->cols = cols; b->B = (uint32_t)tks_bscsr_packet_size(W);
    b->chunk_cap = 512;
    // FIXED_WIDTH <= 22: value + column fit one word -> re-encode into the BSX device format (bscsr_topk.cuh);
    // TKS_BSCSR_VERBATIM=1 keeps the reference's words (the path the wider formats always take)
    b->bsx = (W + 10 <= 32) && !(std::getenv("TKS_BSCSR_VERBATIM") && std::atoi(std::getenv("TKS_BSCSR_VERBATIM")) != 0);
    const bool drift_free = h->cfg.fixed_drift_free != 0;
    if (drift_free) b->chunk_cap = 256;   // up to B rows can finish per packet: keeps the 12-bit in-chunk row offset in range
    // measurement knobs: packets per work unit (multiple of 32) and the divisor of the tail units
    if (const char *e = std::getenv("TKS_BSCSR_CHUNK")) {
        const uint32_t v = (uint32_t)std::atoi(e);
        if (v >= 64 && v <= 4096 && v % 32 == 0) b->chunk_cap = v;
    }
    if (const char *e = std::getenv("TKS_BSCSR_TAIL_DIV")) {
        const uint32_t v = (uint32_t)std::atoi(e);
        if (v >= 1 && v <= 8 && (b->chunk_cap / v) % 32 == 0) b->tail_div = v;
    }
    if (drift_free && (!b->bsx || LFR < 2))
        return h->fail(TKS_EINVAL, "fixed_drift_free needs fixed_width <= 22 (the re-encoded device format) and limited_finished_rows >= 2");
    *out = b;
    return TKS_OK;
}

static int bscsr_finish_upload(Handle *h, BscsrState *b, uint32_t cols, uint64_t total);

int bscsr_upload(Handle *h, uint32_t cols, uint32_t partitions, const uint64_t *packets_per_part,
                 const void *const *packets, const uint32_t *first_row, const uint64_t *nnz_per_part) {
    if (!packets_per_part || !packets || !first_row) return h->fail(TKS_EINVAL, "null argument");
    BscsrState *b = nullptr;
    int rc0 = bscsr_begin_upload(h, cols, partitions, &b);
    if (rc0) return rc0;
    const int W = h->cfg.fixed_width, LFR = h->cfg.limited_finished_rows;
    const int B = (int)b->B;
    b->first_row.assign(first_row, first_row + partitions);
    uint64_t total = 0;
    for (uint32_t p = 0; p < partitions; p++) {
        if (packets_per_part[p] == 0 || !packets[p]) return h->fail(TKS_EINVAL, "partition %u has no packets", p);
        total += packets_per_part[p];
        if (nnz_per_part) b->total_nnz += nnz_per_part[p];
    }
    if (total > 0xFFFFFFF0ull) return h->fail(TKS_EINVAL, "more than 2^32 packets on one device");
    b->total_packets = total;
    const bool drift_free = h->cfg.fixed_drift_free != 0;
    std::vector<std::vector<uint32_t>> enc(b->bsx ? partitions : 0);
    auto field = [](const uint8_t *pk72, int pos, int len) -> uint32_t {
        uint64_t v;
        std::memcpy(&v, pk72 + pos / 8, 8);
        return (uint32_t)((v >> (pos % 8)) & ((1ull << len) - 1ull));
    };

    // ---- chunk tables (host, once per matrix): row counter and carry look-back at every chunk start ----
    std::vector<uint32_t> c_first, c_count, c_local0, c_row_in, c_look, c_part, part_begin(partitions + 1, 0);
    std::vector<uint32_t> s_first, s_count, s_local0, s_look, s_part, s_part_begin(partitions + 1, 0), sample_end(partitions, 0);
    uint64_t goff = 0;
    const uint64_t tail_begin = total - total / 10u;   // last 10 % of the stream
    for (uint32_t p = 0; p < partitions; p++) {
        part_begin[p] = (uint32_t)c_first.size();
        s_part_begin[p] = (uint32_t)s_first.size();
        const uint8_t *pk = static_cast<const uint8_t *>(packets[p]);
        const uint64_t np = packets_per_part[p];
        uint32_t last_row = 0, chunk_row_in = 0, n_pieces_here = 0;
        uint64_t next_chunk = 0, next_piece = 0;
        if (b->bsx) enc[p].assign((size_t)np * 16, 0u);
        std::vector<uint8_t> keepflag(np);   // packet passes the carried partial sum through (n == 1 && !new)
        for (uint64_t i = 0; i < np; i++) {
            uint64_t w0;
            std::memcpy(&w0, pk + i * 64, 8);
            const uint32_t xf = pk[i * 64 + 63] >> 7;
            // validation: cumulative ends are non-decreasing, start >= 1, end <= B
            uint32_t prev = 0;
            for (int s = 0; s < B; s++) {
                uint32_t xs = (uint32_t)((w0 >> (4 * s)) & 0xF);
                if (s >= 16) break;   // B <= 15 for W >= 20; (4-bit fields of packets with B == 16 spill into word 1)
                if (xs < prev || xs > (uint32_t)B || (s == 0 && xs == 0))
                    return h->fail(TKS_EINVAL, "malformed packet %llu of partition %u (segment ends not in 1..B / decreasing)",
                                   (unsigned long long)i, p);
                prev = xs;
            }
            uint32_t n = 0, pe = 0;
            for (int s = 0; s < LFR; s++) { uint32_t xs = (uint32_t)((w0 >> (4 * s)) & 0xF); n += (xs != pe); pe = xs; }
            const uint32_t nw = (i != 0) ? xf : 0u;
            // drift-free mode (not the reference, see topkspmv.h): a packet with more than LFR row segments keeps its
            // first LFR-1 segments, has the values of the rows that finish without a lane zeroed, and ends with ONE
            // segment that spans them and the true last segment -- so the unchanged kernel carries the true partial sum;
            // its row counter advances by the true number of finished rows
            uint32_t nseg = 0, zero_from = 0, zero_to = 0;
            uint64_t xw = w0;
            if (drift_free) {
                uint32_t st = 0, ends[16];
                for (int s = 0; s < B; s++) {
                    uint32_t xs = (uint32_t)((w0 >> (4 * s)) & 0xF);
                    if (xs != st) ends[nseg++] = xs;
                    st = xs;
                }
                if (nseg > (uint32_t)LFR) {
                    zero_from = ends[LFR - 2];        // end of segment LFR-2 = start of the first row without a lane
                    zero_to = ends[nseg - 2];         // start of the true last segment
                    xw = 0;
                    for (int s = 0; s < 16; s++) {
                        const uint32_t e = (s < LFR - 1) ? ends[s] : ends[nseg - 1];
                        xw |= (uint64_t)e << (4 * s);
                    }
                }
            }
            const uint32_t rows_done = (drift_free && nseg > (uint32_t)LFR) ? nseg : n;
            if (i == next_chunk) {
                uint32_t L = 0;
                if (i > 0) { L = 1; while (i - L > 0 && keepflag[i - L]) L++; }
                // look-back + chunk = a whole number of 32-packet warp iterations (no extra iteration for the look-back)
                // the chunks processed last are four times smaller: the persistent warps then run dry within a
                // few iterations of each other instead of up to a whole 512-packet chunk apart
                const uint32_t cap_here = (goff + i >= tail_begin) ? b->chunk_cap / b->tail_div : b->chunk_cap;
                const uint64_t cnt = std::min<uint64_t>(cap_here - (L % 32u), np - i);
                c_first.push_back((uint32_t)(goff + i));
                c_count.push_back((uint32_t)cnt);
                c_local0.push_back((uint32_t)i);
                c_row_in.push_back(last_row);
                c_look.push_back(L);
                c_part.push_back(p);
                next_chunk = i + cnt;
                chunk_row_in = last_row;
            }
            if (b->bsx) {
                uint8_t pk72[72] = {0};
                std::memcpy(pk72, pk + i * 64, 64);
                uint32_t *wout = enc[p].data() + (size_t)i * 16;
                for (int j = 0; j < B; j++) {
                    const uint32_t v = ((uint32_t)j >= zero_from && (uint32_t)j < zero_to) ? 0u : field(pk72, 14 * B + W * j, W);
                    wout[j] = (v << (32 - W)) | field(pk72, 4 * B + 10 * j, 10);
                }
                const uint32_t rel = last_row - chunk_row_in + nw;
                if (rel > 0xFFFu) return h->fail(TKS_EINVAL, "internal: row offset inside a chunk exceeds 12 bits");
                wout[15] = (uint32_t)(xw & 0xFFFFu) | ((nw | (n << 1) | (rel << 4)) << 16);
            }
            if (i == next_piece && i < kBsSamplePackets && n_pieces_here < kBsMaxPieces) {
                uint32_t L = 0;
                if (i > 0) { L = 1; while (i - L > 0 && keepflag[i - L]) L++; }
                // look-back + piece = two warp iterations
                const uint64_t lim = std::min<uint64_t>(np, kBsSamplePackets);
                const uint64_t cnt = std::min<uint64_t>(kBsSamplePiece - (L % 32u), lim - i);
                s_first.push_back((uint32_t)(goff + i));
                s_count.push_back((uint32_t)cnt);
                s_local0.push_back((uint32_t)i);
                s_look.push_back(L);
                s_part.push_back(p);
                next_piece = i + cnt;
                n_pieces_here++;
                sample_end[p] = (uint32_t)next_piece;
            }
            last_row += rows_done + nw - 1u;
            keepflag[i] = (n == 1 && nw == 0) || (n == 0 && nw != 0);
        }
        goff += np;
    }
    part_begin[partitions] = (uint32_t)c_first.size();
    s_part_begin[partitions] = (uint32_t)s_first.size();
    b->n_chunks = (uint32_t)c_first.size();
    b->n_pieces = (uint32_t)s_first.size();

    // ---- device memory ----
    TKS_CUDA(h, cudaSetDevice(h->device));
    TKS_CUDA(h, cudaMalloc(&b->d_packets, total * 64));
    goff = 0;
    for (uint32_t p = 0; p < partitions; p++) {
        const void *src = b->bsx ? static_cast<const void *>(enc[p].data()) : packets[p];
        TKS_CUDA(h, cudaMemcpy(b->d_packets + goff * 64, src, packets_per_part[p] * 64, cudaMemcpyHostToDevice));
        goff += packets_per_part[p];
        if (b->bsx) std::vector<uint32_t>().swap(enc[p]);
    }
    auto up = [&](uint32_t **d, const std::vector<uint32_t> &v) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, std::max<size_t>(1, v.size()) * 4);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*d, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
    };
    TKS_CUDA(h, up(&b->d_chunk_first, c_first));
    TKS_CUDA(h, up(&b->d_chunk_count, c_count));
    TKS_CUDA(h, up(&b->d_chunk_local0, c_local0));
    TKS_CUDA(h, up(&b->d_chunk_row_in, c_row_in));
    TKS_CUDA(h, up(&b->d_chunk_lookback, c_look));
    TKS_CUDA(h, up(&b->d_part_chunk_begin, part_begin));
    TKS_CUDA(h, up(&b->d_chunk_part, c_part));
    TKS_CUDA(h, up(&b->d_s_first, s_first));
    TKS_CUDA(h, up(&b->d_s_count, s_count));
    TKS_CUDA(h, up(&b->d_s_local0, s_local0));
    TKS_CUDA(h, up(&b->d_s_lookback, s_look));
    TKS_CUDA(h, up(&b->d_s_part, s_part));
    TKS_CUDA(h, up(&b->d_s_part_begin, s_part_begin));
    TKS_CUDA(h, up(&b->d_sample_end, sample_end));
    return bscsr_finish_upload(h, b, cols, total);
}

// Common second half: per-query scratch, logs, result words, statistics.
static int bscsr_finish_upload(Handle *h, BscsrState *b, uint32_t cols, uint64_t total) {
    const int LFR = h->cfg.limited_finished_rows, Kp = h->cfg.local_k, B = (int)b->B;
    const uint32_t partitions = b->P;
    TKS_CUDA(h, cudaMalloc(&b->d_piece_top, (size_t)b->n_pieces * LFR * 32 * 4));
    TKS_CUDA(h, cudaMalloc(&b->d_ticket, partitions * 4));
    TKS_CUDA(h, cudaMemset(b->d_ticket, 0, partitions * 4));
    TKS_CUDA(h, cudaMalloc(&b->d_theta_seed, (size_t)partitions * LFR * 4));
    TKS_CUDA(h, cudaMemset(b->d_theta_seed, 0, (size_t)partitions * LFR * 4));
    const size_t nlog = (size_t)b->n_chunks * LFR;
    TKS_CUDA(h, cudaMalloc(&b->logs.val, nlog * b->chunk_cap * 4));
    TKS_CUDA(h, cudaMalloc(&b->logs.row, nlog * b->chunk_cap * 4));
    TKS_CUDA(h, cudaMalloc(&b->logs.cnt, nlog * 4));
    TKS_CUDA(h, cudaMalloc(&b->logs.p0, nlog * 4));
    TKS_CUDA(h, cudaMemset(b->logs.p0, 0, nlog * 4));
    TKS_CUDA(h, cudaMalloc(&b->d_xq, 1024 * 4));
    TKS_CUDA(h, cudaMallocHost(&b->h_xq, 1024 * 4));
    TKS_CUDA(h, cudaMalloc(&b->d_counter, 16));   // [0] chunk scheduler, [1] log entries of the last run
    TKS_CUDA(h, cudaMemset(b->d_counter, 0, 16));
    const size_t nres = (size_t)partitions * Kp * 16;
    // index words and value words in one block: one device-to-host copy per run
    TKS_CUDA(h, cudaMalloc(&b->d_res_idx, 2 * nres * 4));
    b->d_res_val = b->d_res_idx + nres;
    TKS_CUDA(h, cudaMemset(b->d_res_idx, 0, 2 * nres * 4));   // positions >= LFR stay 0 (.cpp:100-110)
    TKS_CUDA(h, cudaMallocHost(&b->h_res_idx, 2 * nres * 4));
    b->h_res_val = b->h_res_idx + nres;
    if (const char *v = std::getenv("TKS_BSCSR_VARIANT")) b->variant = std::atoi(v);

    h->rows = 0; h->cols = cols; h->nnz = b->total_nnz;
    h->have_matrix = true;
    h->stats.rows = 0; h->stats.cols = cols; h->stats.nnz = b->total_nnz; h->stats.packets = total;
    h->stats.device_bytes = total * 64;
    // 64 * sum ceil(nnz_p / B) + 64 * ceil(C / B) + P * Kp * 128
    h->stats.algorithmic_bytes = 64ull * total + 64ull * ((cols + B - 1) / B) + (uint64_t)partitions * Kp * 128ull;
    h->stats.launches_per_run = 3;
    return TKS_OK;
}

// ---------------------------------------------------------------------------------------------
// GPU-side packer: partitioning + packets + device tables from row-sorted COO.
// ---------------------------------------------------------------------------------------------
namespace {

template <int W>
void launch_pack(bool bsx, uint32_t grid, cudaStream_t s, const uint32_t *row, const uint32_t *col, const uint32_t *val32,
                 PackParts parts, uint64_t total, int LFR, int drift_free, uint8_t *packets, uint32_t *advance, uint8_t *keep) {
    if constexpr (W + 10 <= 32) {
        if (bsx) {
            bscsr_pack_kernel<W, true><<<grid, 128, 0, s>>>(row, col, val32, parts, total, LFR, drift_free, packets, advance, keep);
            return;
        }
    }
    bscsr_pack_kernel<W, false><<<grid, 128, 0, s>>>(row, col, val32, parts, total, LFR, drift_free, packets, advance, keep);
}

struct DevFree {   // frees the temporaries of bscsr_upload_coo on every exit path
    std::vector<void *> ptrs;
    ~DevFree() { for (void *p : ptrs) cudaFree(p); }
    template <typename T> cudaError_t alloc(T **p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

}  // namespace

int bscsr_upload_coo(Handle *h, const uint32_t *row, const uint32_t *col, const uint32_t *val32, uint64_t nnz,
                     uint32_t num_rows, uint32_t cols, bool arrays_on_device) {
    if (!row || !col || !val32) return h->fail(TKS_EINVAL, "null argument");
    if (nnz == 0 || num_rows == 0) return h->fail(TKS_EINVAL, "empty matrix");
    BscsrState *b = nullptr;
    const uint32_t P = (uint32_t)h->cfg.partitions;
    int rc = bscsr_begin_upload(h, cols, P, &b);
    if (rc) return rc;
    const int W = h->cfg.fixed_width, LFR = h->cfg.limited_finished_rows;
    const int B = (int)b->B;
    const int drift_free = h->cfg.fixed_drift_free != 0;
    TKS_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    DevFree tmp;
    const uint32_t *d_row = row, *d_col = col, *d_val = val32;
    if (!arrays_on_device) {
        uint32_t *r = nullptr, *c = nullptr, *v = nullptr;
        TKS_CUDA(h, tmp.alloc(&r, nnz * 4)); TKS_CUDA(h, tmp.alloc(&c, nnz * 4)); TKS_CUDA(h, tmp.alloc(&v, nnz * 4));
        TKS_CUDA(h, cudaMemcpyAsync(r, row, nnz * 4, cudaMemcpyHostToDevice, s));
        TKS_CUDA(h, cudaMemcpyAsync(c, col, nnz * 4, cudaMemcpyHostToDevice, s));
        TKS_CUDA(h, cudaMemcpyAsync(v, val32, nnz * 4, cudaMemcpyHostToDevice, s));
        d_row = r; d_col = c; d_val = v;
    }
    // input checks + row partitioning (host_spmv_bscsr.cpp:136-150)
    uint32_t *d_err = nullptr;
    uint64_t *d_nnz_start = nullptr, *d_pkt_start = nullptr;
    TKS_CUDA(h, tmp.alloc(&d_err, 4));
    TKS_CUDA(h, tmp.alloc(&d_nnz_start, (P + 1) * 8));
    TKS_CUDA(h, tmp.alloc(&d_pkt_start, (P + 1) * 8));
    TKS_CUDA(h, cudaMemsetAsync(d_err, 0, 4, s));
    bscsr_pack_check_kernel<<<h->num_sms * 8, 256, 0, s>>>(d_row, d_col, nnz, num_rows, cols, d_err);
    const uint32_t rpp = (num_rows + P - 1) / P;
    bscsr_partition_kernel<<<(P + 1 + 127) / 128, 128, 0, s>>>(d_row, nnz, rpp, P, d_nnz_start);
    std::vector<uint64_t> nnz_start(P + 1), pkt_start(P + 1, 0);
    uint32_t herr = 0;
    TKS_CUDA(h, cudaMemcpyAsync(nnz_start.data(), d_nnz_start, (P + 1) * 8, cudaMemcpyDeviceToHost, s));
    TKS_CUDA(h, cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, s));
    TKS_CUDA(h, cudaStreamSynchronize(s));
    if (herr)
        return h->fail(TKS_EINVAL, "invalid COO:%s%s%s", (herr & kPackErrUnsorted) ? " rows are not sorted;" : "",
                       (herr & kPackErrRowRange) ? " row index beyond num_rows;" : "",
                       (herr & kPackErrColRange) ? " column index >= cols;" : "");
    for (uint32_t p = 0; p < P; p++) {
        const uint64_t n = nnz_start[p + 1] - nnz_start[p];
        if (n == 0)
            return h->fail(TKS_EINVAL, "partition %u has no non-zeros (the reference requires every row range to be populated)", p);
        pkt_start[p + 1] = pkt_start[p] + (n + (uint64_t)B - 1) / (uint64_t)B;
    }
    const uint64_t total = pkt_start[P];
    if (total > 0xFFFFFFF0ull) return h->fail(TKS_EINVAL, "more than 2^32 packets on one device");
    b->total_packets = total;
    b->total_nnz = nnz;
    TKS_CUDA(h, cudaMemcpyAsync(d_pkt_start, pkt_start.data(), (P + 1) * 8, cudaMemcpyHostToDevice, s));
    // first row of every partition (host:145): one 4-byte read each
    b->first_row.assign(P, 0);
    for (uint32_t p = 0; p < P; p++)
        TKS_CUDA(h, cudaMemcpyAsync(&b->first_row[p], d_row + nnz_start[p], 4, cudaMemcpyDeviceToHost, s));

    // packets + per-packet row-counter advance and carry flag
    uint32_t *d_adv = nullptr;
    uint8_t *d_keep = nullptr;
    uint64_t *d_rowsum = nullptr;
    TKS_CUDA(h, cudaMalloc(&b->d_packets, total * 64));
    TKS_CUDA(h, tmp.alloc(&d_adv, total * 4));
    TKS_CUDA(h, tmp.alloc(&d_keep, total));
    TKS_CUDA(h, tmp.alloc(&d_rowsum, (total + 1) * 8));
    const PackParts parts{d_nnz_start, d_pkt_start, P};
    const uint32_t pgrid = (uint32_t)((total + 127) / 128);
    switch (W) {
        case 20: launch_pack<20>(b->bsx, pgrid, s, d_row, d_col, d_val, parts, total, LFR, drift_free, b->d_packets, d_adv, d_keep); break;
        case 21: launch_pack<21>(b->bsx, pgrid, s, d_row, d_col, d_val, parts, total, LFR, drift_free, b->d_packets, d_adv, d_keep); break;
        case 25: launch_pack<25>(b->bsx, pgrid, s, d_row, d_col, d_val, parts, total, LFR, drift_free, b->d_packets, d_adv, d_keep); break;
        case 26: launch_pack<26>(b->bsx, pgrid, s, d_row, d_col, d_val, parts, total, LFR, drift_free, b->d_packets, d_adv, d_keep); break;
        default: launch_pack<32>(b->bsx, pgrid, s, d_row, d_col, d_val, parts, total, LFR, drift_free, b->d_packets, d_adv, d_keep); break;
    }
    TKS_CUDA(h, cudaGetLastError());
    rc = device_scan_u32(h, d_adv, total, d_rowsum);   // synchronises the stream
    if (rc) return rc;

    // chunk and sample-piece tables: count, then fill
    uint32_t *d_nch = nullptr, *d_npc = nullptr, *d_cbeg = nullptr, *d_pbeg = nullptr;
    TKS_CUDA(h, tmp.alloc(&d_nch, P * 4)); TKS_CUDA(h, tmp.alloc(&d_npc, P * 4));
    TKS_CUDA(h, cudaMalloc(&b->d_part_chunk_begin, (P + 1) * 4));
    TKS_CUDA(h, cudaMalloc(&b->d_s_part_begin, (P + 1) * 4));
    d_cbeg = b->d_part_chunk_begin; d_pbeg = b->d_s_part_begin;
    WalkOut wo{};
    wo.n_chunks = d_nch; wo.n_pieces = d_npc;
    bscsr_chunk_walk_kernel<<<(P + 31) / 32, 32, 0, s>>>(parts, d_keep, d_rowsum, total, b->chunk_cap, b->tail_div, wo);
    std::vector<uint32_t> nch(P), npc(P), cbeg(P + 1, 0), pbeg(P + 1, 0);
    TKS_CUDA(h, cudaMemcpyAsync(nch.data(), d_nch, P * 4, cudaMemcpyDeviceToHost, s));
    TKS_CUDA(h, cudaMemcpyAsync(npc.data(), d_npc, P * 4, cudaMemcpyDeviceToHost, s));
    TKS_CUDA(h, cudaStreamSynchronize(s));
    for (uint32_t p = 0; p < P; p++) { cbeg[p + 1] = cbeg[p] + nch[p]; pbeg[p + 1] = pbeg[p] + npc[p]; }
    b->n_chunks = cbeg[P];
    b->n_pieces = pbeg[P];
    TKS_CUDA(h, cudaMemcpyAsync(d_cbeg, cbeg.data(), (P + 1) * 4, cudaMemcpyHostToDevice, s));
    TKS_CUDA(h, cudaMemcpyAsync(d_pbeg, pbeg.data(), (P + 1) * 4, cudaMemcpyHostToDevice, s));
    const size_t nc4 = std::max<size_t>(1, b->n_chunks) * 4, np4 = std::max<size_t>(1, b->n_pieces) * 4;
    TKS_CUDA(h, cudaMalloc(&b->d_chunk_first, nc4)); TKS_CUDA(h, cudaMalloc(&b->d_chunk_count, nc4));
    TKS_CUDA(h, cudaMalloc(&b->d_chunk_local0, nc4)); TKS_CUDA(h, cudaMalloc(&b->d_chunk_row_in, nc4));
    TKS_CUDA(h, cudaMalloc(&b->d_chunk_lookback, nc4)); TKS_CUDA(h, cudaMalloc(&b->d_chunk_part, nc4));
    TKS_CUDA(h, cudaMalloc(&b->d_s_first, np4)); TKS_CUDA(h, cudaMalloc(&b->d_s_count, np4));
    TKS_CUDA(h, cudaMalloc(&b->d_s_local0, np4)); TKS_CUDA(h, cudaMalloc(&b->d_s_lookback, np4));
    TKS_CUDA(h, cudaMalloc(&b->d_s_part, np4));
    TKS_CUDA(h, cudaMalloc(&b->d_sample_end, P * 4));
    wo.first = b->d_chunk_first; wo.count = b->d_chunk_count; wo.local0 = b->d_chunk_local0; wo.row_in = b->d_chunk_row_in;
    wo.look = b->d_chunk_lookback; wo.part = b->d_chunk_part;
    wo.s_first = b->d_s_first; wo.s_count = b->d_s_count; wo.s_local0 = b->d_s_local0; wo.s_look = b->d_s_lookback;
    wo.s_part = b->d_s_part; wo.sample_end = b->d_sample_end;
    wo.chunk_begin = d_cbeg; wo.piece_begin = d_pbeg;
    bscsr_chunk_walk_kernel<<<(P + 31) / 32, 32, 0, s>>>(parts, d_keep, d_rowsum, total, b->chunk_cap, b->tail_div, wo);
    if (b->bsx) {
        TKS_CUDA(h, cudaMemsetAsync(d_err, 0, 4, s));
        bscsr_patch_rel_kernel<<<(b->n_chunks * 32u + 255u) / 256u, 256, 0, s>>>(
            b->d_packets, b->d_chunk_first, b->d_chunk_count, b->d_chunk_part, b->d_chunk_row_in, b->n_chunks, d_rowsum,
            d_pkt_start, d_err);
        TKS_CUDA(h, cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, s));
    }
    TKS_CUDA(h, cudaStreamSynchronize(s));
    TKS_CUDA(h, cudaGetLastError());
    if (herr & kPackErrRel) return h->fail(TKS_EINVAL, "internal: row offset inside a chunk exceeds 12 bits");
    return bscsr_finish_upload(h, b, cols, total);
}

static void bscsr_transform_query(const BscsrState *b, int W, const uint32_t *vec32, uint32_t *xq);

int bscsr_set_query(Handle *h, const uint32_t *vec32_host, const uint32_t *vec32_dev, cudaStream_t s) {
    BscsrState *b = h->bs;
    if (!b) return h->fail(TKS_ESTATE, "no packets uploaded");
    const int W = h->cfg.fixed_width;
    if (!vec32_host) {
        switch (W) {
            case 20: bscsr_query_kernel<20><<<4, 256, 0, s>>>(vec32_dev, b->cols, b->d_xq); break;
            case 21: bscsr_query_kernel<21><<<4, 256, 0, s>>>(vec32_dev, b->cols, b->d_xq); break;
            case 25: bscsr_query_kernel<25><<<4, 256, 0, s>>>(vec32_dev, b->cols, b->d_xq); break;
            case 26: bscsr_query_kernel<26><<<4, 256, 0, s>>>(vec32_dev, b->cols, b->d_xq); break;
            default: bscsr_query_kernel<32><<<4, 256, 0, s>>>(vec32_dev, b->cols, b->d_xq); break;
        }
        TKS_CUDA(h, cudaGetLastError());
        b->have_query = true;
        return TKS_OK;
    }
    // kernel vec load (.cpp:127-137): W-bit truncation of the 32-bit word; pre-shifted by one for the
    // umulhi product (see bscsr_stream_kernel); columns >= cols read 0 like the zero-initialised URAM
    if (!b->ev_query) TKS_CUDA(h, cudaEventCreateWithFlags(&b->ev_query, cudaEventDisableTiming));
    TKS_CUDA(h, cudaEventSynchronize(b->ev_query));   // h_xq (pinned staging) may still feed the previous copy
    bscsr_transform_query(b, W, vec32_host, b->h_xq);
    TKS_CUDA(h, cudaMemcpyAsync(b->d_xq, b->h_xq, 1024 * 4, cudaMemcpyHostToDevice, s));
    TKS_CUDA(h, cudaEventRecord(b->ev_query, s));
    b->have_query = true;
    return TKS_OK;
}

static int bscsr_dispatch(Handle *h, BscsrState *b, cudaStream_t s, const BsRun &r) {
    BscsrChunks m{b->d_chunk_first, b->d_chunk_count, b->d_chunk_local0, b->d_chunk_row_in, b->d_chunk_lookback,
                  b->d_chunk_part, b->n_chunks, b->chunk_cap};
    int rc;
    switch (h->cfg.fixed_width) {
        case 20: rc = dispatch_lfr<20>(h, b, m, s, r); break;
        case 21: rc = dispatch_lfr<21>(h, b, m, s, r); break;
        case 25: rc = dispatch_lfr<25>(h, b, m, s, r); break;
        case 26: rc = dispatch_lfr<26>(h, b, m, s, r); break;
        case 32: rc = dispatch_lfr<32>(h, b, m, s, r); break;
        default: return h->fail(TKS_EINVAL, "fixed_width not instantiated");
    }
    if (rc) return rc;
    TKS_CUDA(h, cudaGetLastError());
    return TKS_OK;
}

static int bscsr_pipe_drain(Handle *h, BscsrState *b) {
    for (int i = 0; i < BscsrState::kSlots; i++) {
        if (!b->p_busy[i]) continue;
        TKS_CUDA(h, cudaEventSynchronize(b->p_ev_done[i]));
        b->p_busy[i] = false;
    }
    return TKS_OK;
}

int bscsr_launch(Handle *h, cudaStream_t s) {
    BscsrState *b = h->bs;
    if (!b) return h->fail(TKS_ESTATE, "no packets uploaded");
    if (!b->have_query) return h->fail(TKS_ESTATE, "no query set");
    int rcd = bscsr_pipe_drain(h, b);   // the logs and the chunk scheduler are shared with pipelined submits
    if (rcd) return rcd;
    if (s != h->stream && b->ev_query) TKS_CUDA(h, cudaStreamWaitEvent(s, b->ev_query, 0));
    const BsRun r{b->d_xq, b->d_theta_seed, b->d_res_idx, b->d_res_val, nullptr, nullptr};
    int rc = bscsr_dispatch(h, b, s, r);
    if (rc) return rc;
    b->have_words = false;
    return TKS_OK;
}

// W-bit truncation of the raw query words, pre-shifted for the umulhi product (kernel vec load, .cpp:127-137)
static void bscsr_transform_query(const BscsrState *b, int W, const uint32_t *vec32, uint32_t *xq) {
    for (uint32_t c = 0; c < 1024; c++) {
        const uint32_t q = (c < b->cols) ? (vec32[c] >> (32 - W)) : 0u;
        xq[c] = (W == 32) ? q : (q << 1);
    }
}

static int bscsr_pipe_init(Handle *h, BscsrState *b) {
    if (b->p_sample_stream) return TKS_OK;
    int lo = 0, hi = 0;
    TKS_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    TKS_CUDA(h, cudaStreamCreateWithPriority(&b->p_sample_stream, cudaStreamNonBlocking, hi));
    const size_t nres = (size_t)b->P * h->cfg.local_k * 16;
    for (int i = 0; i < BscsrState::kSlots; i++) {
        TKS_CUDA(h, cudaEventCreateWithFlags(&b->p_ev_sample[i], cudaEventDisableTiming));
        TKS_CUDA(h, cudaEventCreateWithFlags(&b->p_ev_done[i], cudaEventDisableTiming));
        TKS_CUDA(h, cudaMalloc(&b->p_d_xq[i], 1024 * 4));
        TKS_CUDA(h, cudaMallocHost(&b->p_h_xq[i], 1024 * 4));
        TKS_CUDA(h, cudaMalloc(&b->p_d_theta[i], (size_t)b->P * h->cfg.limited_finished_rows * 4));
        TKS_CUDA(h, cudaMemset(b->p_d_theta[i], 0, (size_t)b->P * h->cfg.limited_finished_rows * 4));
        TKS_CUDA(h, cudaMallocHost(&b->p_h_words[i], 2 * nres * 4));
        std::memset(b->p_h_words[i], 0, 2 * nres * 4);   // positions >= LFR stay 0 (.cpp:100-110)
    }
    return TKS_OK;
}

// One query into the two-slot pipeline: transform + H2D + sample on the sample stream (they overlap the previous
// query's stream and replay kernels), then stream + replay on stream `s`; for a host query the replay kernel stores the
// result words into the slot's pinned host block and bscsr_fetch_ticket() waits for that query alone and merges.
int bscsr_submit(Handle *h, const uint32_t *vec32_host, const uint32_t *vec32_dev, uint32_t k, cudaStream_t s,
                 bool query_ready, uint64_t *ticket) {
    BscsrState *b = h->bs;
    if (!b) return h->fail(TKS_ESTATE, "no packets uploaded");
    int rc = bscsr_pipe_init(h, b);
    if (rc) return rc;
    const uint64_t seq = b->p_seq + 1;
    const int slot = (int)(seq % BscsrState::kSlots);
    if (b->p_busy[slot]) {
        TKS_CUDA(h, cudaEventSynchronize(b->p_ev_done[slot]));
        b->p_busy[slot] = false;
    }
    const size_t nres = (size_t)b->P * h->cfg.local_k * 16;
    uint32_t *res_idx = b->d_res_idx, *res_val = b->d_res_val;
    if (vec32_host) {
        bscsr_transform_query(b, h->cfg.fixed_width, vec32_host, b->p_h_xq[slot]);
        TKS_CUDA(h, cudaMemcpyAsync(b->p_d_xq[slot], b->p_h_xq[slot], 1024 * 4, cudaMemcpyHostToDevice, b->p_sample_stream));
        res_idx = b->p_h_words[slot];            // the replay kernel writes the words straight to pinned host memory
        res_val = b->p_h_words[slot] + nres;
        b->p_ticket[slot] = seq;
    } else {
        // a query already in HBM (raw words): transformed on the sample stream; the result words stay on the device
        if (!query_ready) {
            if (!b->ev_query) TKS_CUDA(h, cudaEventCreateWithFlags(&b->ev_query, cudaEventDisableTiming));
            TKS_CUDA(h, cudaEventRecord(b->ev_query, s));
            TKS_CUDA(h, cudaStreamWaitEvent(b->p_sample_stream, b->ev_query, 0));
        }
        switch (h->cfg.fixed_width) {
            case 20: bscsr_query_kernel<20><<<4, 256, 0, b->p_sample_stream>>>(vec32_dev, b->cols, b->p_d_xq[slot]); break;
            case 21: bscsr_query_kernel<21><<<4, 256, 0, b->p_sample_stream>>>(vec32_dev, b->cols, b->p_d_xq[slot]); break;
            case 25: bscsr_query_kernel<25><<<4, 256, 0, b->p_sample_stream>>>(vec32_dev, b->cols, b->p_d_xq[slot]); break;
            case 26: bscsr_query_kernel<26><<<4, 256, 0, b->p_sample_stream>>>(vec32_dev, b->cols, b->p_d_xq[slot]); break;
            default: bscsr_query_kernel<32><<<4, 256, 0, b->p_sample_stream>>>(vec32_dev, b->cols, b->p_d_xq[slot]); break;
        }
        b->p_ticket[slot] = 0;
    }
    const BsRun r{b->p_d_xq[slot], b->p_d_theta[slot], res_idx, res_val, b->p_sample_stream, b->p_ev_sample[slot]};
    rc = bscsr_dispatch(h, b, s, r);
    if (rc) return rc;
    TKS_CUDA(h, cudaEventRecord(b->p_ev_done[slot], s));
    b->p_busy[slot] = true;
    b->p_k[slot] = k;
    b->p_seq = seq;
    b->have_words = false;
    h->last_k = k;
    if (ticket) *ticket = seq;
    return TKS_OK;
}

int bscsr_submit_host(Handle *h, const uint32_t *vec32, uint32_t k, uint64_t *ticket) {
    return bscsr_submit(h, vec32, nullptr, k, h->stream, true, ticket);
}

int bscsr_fetch_ticket(Handle *h, uint64_t ticket, uint32_t *idx_out, uint32_t *val_out, uint32_t *count) {
    BscsrState *b = h->bs;
    if (!b || !b->p_sample_stream || ticket == 0 || ticket > b->p_seq) return h->fail(TKS_EINVAL, "unknown ticket");
    const int slot = (int)(ticket % BscsrState::kSlots);
    if (b->p_ticket[slot] != ticket)
        return h->fail(TKS_ESTATE, "the result of ticket %llu is gone: at most %d queries are kept, fetch before submitting further",
                       (unsigned long long)ticket, BscsrState::kSlots);
    TKS_CUDA(h, cudaEventSynchronize(b->p_ev_done[slot]));
    b->p_busy[slot] = false;
    const size_t nres = (size_t)b->P * h->cfg.local_k * 16;
    const uint32_t k = b->p_k[slot];
    std::vector<uint32_t> mi(k, 0u), mv(k, 0u);
    uint32_t n_out = 0;
    // read_result (host_spmv_bscsr.cpp:399-448) + sort_tuples over the slot's words
    if (tks_merge_partition_words(b->P, (uint32_t)h->cfg.local_k, b->B, b->p_h_words[slot], b->p_h_words[slot] + nres,
                                  b->first_row.data(), h->cfg.tie_break, k, mi.data(), mv.data(), &n_out) != TKS_OK)
        return h->fail(TKS_EINVAL, "merge of the partition results failed");
    const uint32_t n = n_out < k ? n_out : k;
    std::memcpy(idx_out, mi.data(), n * 4);
    std::memcpy(val_out, mv.data(), n * 4);
    for (uint32_t i = n; i < k; i++) { idx_out[i] = 0; val_out[i] = 0; }
    if (count) *count = n;
    h->stats.last_candidates = n_out;
    return TKS_OK;
}

int bscsr_fetch(Handle *h) {
    BscsrState *b = h->bs;
    if (!b) return h->fail(TKS_ESTATE, "no packets uploaded");
    const size_t nres = (size_t)b->P * h->cfg.local_k * 16;
    TKS_CUDA(h, cudaMemcpyAsync(b->h_res_idx, b->d_res_idx, 2 * nres * 4, cudaMemcpyDeviceToHost, h->stream));
    TKS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->cfg.profile_kernels) {   // statistics: candidates the stream kernel logged for the replay
        uint32_t logged = 0;
        if (cudaMemcpy(&logged, b->d_counter + 1, 4, cudaMemcpyDeviceToHost) == cudaSuccess) h->stats.logged_candidates = logged;
    }
    b->have_words = true;
    // read_result (host_spmv_bscsr.cpp:399-448) + sort_tuples: the host-side merge of host_api.cpp
    // only the first k of the merged list are ever read (tks_read_result with the run's k): a partial sort is enough
    const uint32_t cap = b->P * (uint32_t)h->cfg.local_k * b->B;
    const uint32_t want = h->last_k && h->last_k < cap ? h->last_k : cap;
    b->merged_idx.assign(want, 0u);
    b->merged_val.assign(want, 0u);
    uint32_t n_out = 0;
    if (tks_merge_partition_words(b->P, (uint32_t)h->cfg.local_k, b->B, b->h_res_idx, b->h_res_val, b->first_row.data(),
                                  h->cfg.tie_break, want, b->merged_idx.data(), b->merged_val.data(), &n_out) != TKS_OK)
        return h->fail(TKS_EINVAL, "merge of the partition results failed");
    const uint32_t have = n_out < want ? n_out : want;
    b->merged_idx.resize(have);
    b->merged_val.resize(have);
    h->stats.last_candidates = n_out;
    return TKS_OK;
}

int bscsr_read_result(Handle *h, uint32_t *idx_out, uint32_t *val_out, uint32_t k, uint32_t *count) {
    BscsrState *b = h->bs;
    if (!b) return h->fail(TKS_ESTATE, "no packets uploaded");
    if (!b->have_words) {
        TKS_CUDA(h, cudaSetDevice(h->device));
        TKS_CUDA(h, cudaDeviceSynchronize());
        int rc = bscsr_fetch(h);
        if (rc) return rc;
    }
    const uint32_t n = (uint32_t)std::min<size_t>(k, b->merged_idx.size());
    std::memcpy(idx_out, b->merged_idx.data(), n * 4);
    std::memcpy(val_out, b->merged_val.data(), n * 4);
    for (uint32_t i = n; i < k; i++) { idx_out[i] = 0; val_out[i] = 0; }
    if (count) *count = n;
    return TKS_OK;
}

// Device address of the last un-pipelined run's result words: partitions x local_k x 16 index words, then as many value
// words (one block), for a device-side all-gather between ranks.
int bscsr_partition_words_device(Handle *h, const uint32_t **d_words, uint32_t *n_words) {
    BscsrState *b = h->bs;
    if (!b) return h->fail(TKS_ESTATE, "no packets uploaded");
    *d_words = b->d_res_idx;
    if (n_words) *n_words = (uint32_t)(2u * (size_t)b->P * h->cfg.local_k * 16u);
    return TKS_OK;
}

int bscsr_read_partition_results(Handle *h, uint32_t *idx_words, uint32_t *val_words) {
    BscsrState *b = h->bs;
    if (!b) return h->fail(TKS_ESTATE, "no packets uploaded");
    if (!b->have_words) {
        TKS_CUDA(h, cudaSetDevice(h->device));
        TKS_CUDA(h, cudaDeviceSynchronize());
        int rc = bscsr_fetch(h);
        if (rc) return rc;
    }
    const size_t nres = (size_t)b->P * h->cfg.local_k * 16;
    if (idx_words) std::memcpy(idx_words, b->h_res_idx, nres * 4);
    if (val_words) std::memcpy(val_words, b->h_res_val, nres * 4);
    return TKS_OK;
}

// FNV-1a digests of the resident matrix state, in a fixed order (packets, the six chunk tables, the partition
// chunk offsets, the five sample tables, their offsets, sample_end, first_row): lets a test assert that two upload
// paths leave byte-identical device state without shipping it back through the ABI.
int bscsr_state_digest(Handle *h, uint64_t *digest, uint32_t n) {
    BscsrState *b = h->bs;
    if (!b || !h->have_matrix) return h->fail(TKS_ESTATE, "no BS-CSR matrix resident");
    if (n < 16) return h->fail(TKS_EINVAL, "digest array needs 16 entries");
    TKS_CUDA(h, cudaSetDevice(h->device));
    auto fnv = [&](const void *dptr, size_t bytes, uint64_t *out) -> cudaError_t {
        std::vector<uint8_t> buf(bytes);
        cudaError_t e = cudaMemcpy(buf.data(), dptr, bytes, cudaMemcpyDeviceToHost);
        uint64_t hsh = 1469598103934665603ull;
        for (uint8_t c : buf) { hsh ^= c; hsh *= 1099511628211ull; }
        *out = hsh;
        return e;
    };
    const size_t nc = (size_t)b->n_chunks * 4, np = (size_t)b->n_pieces * 4;
    TKS_CUDA(h, fnv(b->d_packets, b->total_packets * 64, &digest[0]));
    TKS_CUDA(h, fnv(b->d_chunk_first, nc, &digest[1]));
    TKS_CUDA(h, fnv(b->d_chunk_count, nc, &digest[2]));
    TKS_CUDA(h, fnv(b->d_chunk_local0, nc, &digest[3]));
    TKS_CUDA(h, fnv(b->d_chunk_row_in, nc, &digest[4]));
    TKS_CUDA(h, fnv(b->d_chunk_lookback, nc, &digest[5]));
    TKS_CUDA(h, fnv(b->d_chunk_part, nc, &digest[6]));
    TKS_CUDA(h, fnv(b->d_part_chunk_begin, ((size_t)b->P + 1) * 4, &digest[7]));
    TKS_CUDA(h, fnv(b->d_s_first, np, &digest[8]));
    TKS_CUDA(h, fnv(b->d_s_count, np, &digest[9]));
    TKS_CUDA(h, fnv(b->d_s_local0, np, &digest[10]));
    TKS_CUDA(h, fnv(b->d_s_lookback, np, &digest[11]));
    TKS_CUDA(h, fnv(b->d_s_part, np, &digest[12]));
    TKS_CUDA(h, fnv(b->d_s_part_begin, ((size_t)b->P + 1) * 4, &digest[13]));
    TKS_CUDA(h, fnv(b->d_sample_end, (size_t)b->P * 4, &digest[14]));
    uint64_t hsh = 1469598103934665603ull;
    for (uint32_t r : b->first_row) for (int i = 0; i < 4; i++) { hsh ^= (r >> (8 * i)) & 0xFF; hsh *= 1099511628211ull; }
    digest[15] = hsh ^ ((uint64_t)b->n_chunks << 32) ^ b->n_pieces;
    return TKS_OK;
}

void bscsr_destroy(Handle *h) {
    BscsrState *b = h->bs;
    if (!b) return;
    cudaFree(b->d_packets); cudaFree(b->d_chunk_first); cudaFree(b->d_chunk_count); cudaFree(b->d_chunk_local0);
    cudaFree(b->d_chunk_row_in); cudaFree(b->d_chunk_lookback); cudaFree(b->d_part_chunk_begin);
    cudaFree(b->logs.val); cudaFree(b->logs.row); cudaFree(b->logs.cnt); cudaFree(b->logs.p0);
    cudaFree(b->d_chunk_part); cudaFree(b->d_s_first); cudaFree(b->d_s_count); cudaFree(b->d_s_local0); cudaFree(b->d_s_lookback);
    cudaFree(b->d_s_part); cudaFree(b->d_s_part_begin); cudaFree(b->d_piece_top); cudaFree(b->d_ticket); cudaFree(b->d_theta_seed);
    cudaFree(b->d_xq); cudaFreeHost(b->h_xq); cudaFree(b->d_counter); cudaFree(b->d_sample_end);
    cudaFree(b->d_res_idx); cudaFreeHost(b->h_res_idx);
    if (b->ev_query) cudaEventDestroy(b->ev_query);
    for (int i = 0; i < BscsrState::kSlots; i++) {
        if (b->p_ev_sample[i]) cudaEventDestroy(b->p_ev_sample[i]);
        if (b->p_ev_done[i]) cudaEventDestroy(b->p_ev_done[i]);
        cudaFree(b->p_d_xq[i]); cudaFreeHost(b->p_h_xq[i]); cudaFree(b->p_d_theta[i]); cudaFreeHost(b->p_h_words[i]);
    }
    if (b->p_sample_stream) cudaStreamDestroy(b->p_sample_stream);
    delete b;
    h->bs = nullptr;
}

}  // namespace tks

extern "C" int tks_upload_coo_fixed(tks_handle *h, const uint32_t *row, const uint32_t *col, const uint32_t *val32,
                                    uint64_t nnz, uint32_t num_rows, uint32_t cols) {
    if (!h) return TKS_EINVAL;
    return tks::bscsr_upload_coo(h, row, col, val32, nnz, num_rows, cols, false);
}

extern "C" int tks_upload_coo_fixed_device(tks_handle *h, const uint32_t *d_row, const uint32_t *d_col,
                                           const uint32_t *d_val32, uint64_t nnz, uint32_t num_rows, uint32_t cols) {
    if (!h) return TKS_EINVAL;
    return tks::bscsr_upload_coo(h, d_row, d_col, d_val32, nnz, num_rows, cols, true);
}

extern "C" int tks_bscsr_state_digest(tks_handle *h, uint64_t *digest, uint32_t n) {
    if (!h || !digest) return TKS_EINVAL;
    return tks::bscsr_state_digest(h, digest, n);
}

extern "C" int tks_upload_bscsr(tks_handle *h, uint32_t cols, uint32_t partitions, const uint64_t *packets_per_part,
                                const void *const *packets, const uint32_t *first_row, const uint64_t *nnz_per_part) {
    if (!h) return TKS_EINVAL;
    return tks::bscsr_upload(h, cols, partitions, packets_per_part, packets, first_row, nnz_per_part);
}
