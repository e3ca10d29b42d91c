/*
 * oracle.c -- CPU restatement of the reference's Top-K SpMV algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under approximate-spmv-topk_b200/ may
 * include, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Plain sequential C, no threads, no SIMD: the
 * point is to be obviously equal to the reference, not to be fast.
 *
 * Pinning status
 *   float path  (orc_gold_topk_f32, orc_sort_tuples_f32): PINNED against the
 *       reference's own gold_algorithms.hpp compiled here (oracle/_ref,
 *       tests/test_oracle_vs_ref.py) and the committed goldens.
 *   fixed path  (orc_pack_bscsr, orc_bscsr_partition, orc_read_result):
 *       PINNED against the unmodified reference HLS kernel + host packer
 *       compiled with the ap_fixed/hls_stream/OpenCL shims of oracle/shim
 *       (oracle/_ref/libref_fpga.so, tests/test_oracle_vs_ref.py) when that
 *       library is present; see DESIGN.md "Oracle pinning".
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* Float path                                                                */
/* ------------------------------------------------------------------------ */

/* gold_algorithms.hpp:188-246  spmv_coo_gold_top_k<I=unsigned,V=float>.
 * Streaming replace-min over row-sorted COO; sequential fp32 accumulation in
 * nnz order (separate multiply and add, as g++ emits for x86-64 without FMA);
 * a finished row replaces the current worst slot when `>=`; new worst = first
 * slot holding the minimum (strict `<` scan, :222-227); the final row is
 * flushed without recomputing the worst (:239-244).                          */
ORC_API void orc_gold_topk_f32(const uint32_t *row, const uint32_t *col, const float *val,
                               uint64_t nnz, const float *vec, int k,
                               uint32_t *res_idx, float *res_val) {
    for (int i = 0; i < k; i++) { res_idx[i] = 0; res_val[i] = 0.0f; }
    if (nnz == 0) return;
    uint32_t curr_row = row[0];
    volatile float curr_out = 0.0f;          /* volatile: forbid contraction / reassociation */
    uint32_t worst_idx = 0;
    float worst_val = 0.0f;
    for (uint64_t i = 0; i < nnz; i++) {
        uint32_t r = row[i];
        volatile float contrib = val[i] * vec[col[i]];
        if (r == curr_row) {
            curr_out = curr_out + contrib;
        } else {
            if (curr_out >= worst_val) {
                res_idx[worst_idx] = curr_row;
                res_val[worst_idx] = curr_out;
                uint32_t w = 0;
                float wv = res_val[0];
                for (int j = 0; j < k; j++) {
                    if (res_val[j] < wv) { w = (uint32_t)j; wv = res_val[j]; }
                }
                worst_idx = w;
                worst_val = wv;
            }
            curr_row = r;
            curr_out = contrib;
        }
    }
    if (curr_out >= worst_val) {
        res_idx[worst_idx] = curr_row;
        res_val[worst_idx] = curr_out;
    }
}

/* evaluation_utils.hpp:40-62  sort_tuples: value descending, ties -> HIGHER
 * index first.  Insertion sort keeps it dependency-free (n is k-sized).      */
ORC_API void orc_sort_tuples_f32(uint32_t n, uint32_t *idx, float *val) {
    for (uint32_t i = 1; i < n; i++) {
        uint32_t ki = idx[i]; float kv = val[i];
        int64_t j = (int64_t)i - 1;
        while (j >= 0 && (val[j] < kv || (val[j] == kv && idx[j] < ki))) {
            idx[j + 1] = idx[j]; val[j + 1] = val[j]; j--;
        }
        idx[j + 1] = ki; val[j + 1] = kv;
    }
}

ORC_API void orc_sort_tuples_u32(uint32_t n, uint32_t *idx, uint32_t *val) {
    for (uint32_t i = 1; i < n; i++) {
        uint32_t ki = idx[i]; uint32_t kv = val[i];
        int64_t j = (int64_t)i - 1;
        while (j >= 0 && (val[j] < kv || (val[j] == kv && idx[j] < ki))) {
            idx[j + 1] = idx[j]; val[j + 1] = val[j]; j--;
        }
        idx[j + 1] = ki; val[j + 1] = kv;
    }
}

/* Full fp32 product vector in the same sequential order (used by tests to
 * reason about ties; the sum order equals orc_gold_topk_f32's).              */
ORC_API void orc_spmv_f32(const uint32_t *row, const uint32_t *col, const float *val,
                          uint64_t nnz, const float *vec, float *out, uint32_t num_rows) {
    for (uint32_t r = 0; r < num_rows; r++) out[r] = 0.0f;
    for (uint64_t i = 0; i < nnz; i++) {
        volatile float contrib = val[i] * vec[col[i]];
        volatile float s = out[row[i]] + contrib;
        out[row[i]] = s;
    }
}

/* ------------------------------------------------------------------------ */
/* Fixed-point number format                                                 */
/* ------------------------------------------------------------------------ */

/* fpga_types.hpp:16-23: ap_ufixed<W,1,AP_TRN_ZERO>, default overflow = wrap.
 * Raw value = integer with F = W-1 fractional bits.                          */

/* double -> ap_ufixed<32,1,AP_TRN_ZERO>   (utils.hpp:401 `(T) value`,
 * utils.hpp:242/:264 in create_sample_vector).  Truncate toward zero, wrap.  */
ORC_API uint32_t orc_fx32_from_double(double v) {
    if (!(v > 0.0)) return 0u;                 /* inputs are non-negative; NaN/neg -> 0 */
    double s = floor(v * 2147483648.0);        /* 2^31 */
    s = fmod(s, 4294967296.0);                 /* wrap to 32 bits */
    return (uint32_t)s;
}

/* ap_ufixed<32,1>::to_float() (Vitis ap_fixed_base: round to nearest even),
 * then float -> ap_ufixed<W,1,AP_TRN_ZERO> by truncation, wrap.
 * fpga_utils.hpp:336-338 (write_block_val, !USE_FLOAT branch).               */
ORC_API uint32_t orc_fxW_from_fx32(uint32_t raw32, int W) {
    float f = (float)((double)raw32 / 2147483648.0);   /* exact scale, one RNE rounding */
    double s = floor((double)f * (double)(1ull << (W - 1)));
    uint64_t m = (W == 32) ? 0xFFFFFFFFull : ((1ull << W) - 1ull);
    return (uint32_t)(((uint64_t)s) & m);
}

/* Query word seen by the kernel: 32-bit raw -> W-bit raw by truncation
 * (spmv_bscsr_top_k_multicore.cpp:127-137 `(real_type) curr`).               */
ORC_API uint32_t orc_fxW_from_fx32_trunc(uint32_t raw32, int W) {
    return raw32 >> (32 - W);
}

/* ------------------------------------------------------------------------ */
/* BS-CSR packet builder                                                     */
/* ------------------------------------------------------------------------ */

static void set_bits(uint64_t *w, unsigned lo, unsigned width, uint64_t v) {
    for (unsigned b = 0; b < width; b++) {
        unsigned pos = lo + b;
        uint64_t bit = (v >> b) & 1ull;
        w[pos >> 6] = (w[pos >> 6] & ~(1ull << (pos & 63))) | (bit << (pos & 63));
    }
}

static uint64_t get_bits(const uint64_t *w, unsigned lo, unsigned width) {
    uint64_t v = 0;
    for (unsigned b = 0; b < width; b++) {
        unsigned pos = lo + b;
        v |= ((w[pos >> 6] >> (pos & 63)) & 1ull) << b;
    }
    return v;
}

/* types.hpp:71-72  BSCSR_PACKET_SIZE = (512 - 1) / (W + 10 + 4)              */
ORC_API int orc_packet_size(int W) { return (512 - 1) / (W + 14); }

/* host_spmv_bscsr.cpp:136-150: rows_per_part = ceil(N/P); partition of nnz i
 * is row[i]/rows_per_part.  Fills part_nnz_start[P+1] (prefix into the
 * row-sorted COO), first_row[P], last_row[P], num_packets[P].
 * Returns 0, or -1 if some partition is empty (the reference would read
 * coo_partition[0] of an empty vector there -- undefined; we reject).        */
ORC_API int orc_partition(const uint32_t *row, uint64_t nnz, uint32_t num_rows, int P, int B,
                          uint64_t *part_nnz_start, uint32_t *first_row, uint32_t *last_row,
                          uint64_t *num_packets) {
    uint32_t rpp = (num_rows + (uint32_t)P - 1) / (uint32_t)P;
    uint64_t i = 0;
    for (int p = 0; p < P; p++) {
        part_nnz_start[p] = i;
        while (i < nnz && row[i] / rpp == (uint32_t)p) i++;
        uint64_t n = i - part_nnz_start[p];
        if (n == 0) return -1;
        first_row[p] = row[part_nnz_start[p]];
        last_row[p] = row[i - 1];
        num_packets[p] = (n + (uint64_t)B - 1) / (uint64_t)B;
    }
    part_nnz_start[P] = i;
    return (i == nnz) ? 0 : -2;   /* -2: rows not sorted / row id >= P*rpp */
}

/* host_spmv_bscsr.cpp:189-248 packet_coo_partition + fpga_utils.hpp:307-365.
 * `row/col/val32` point at the partition's first nnz; val32 are raw
 * ap_ufixed<32,1> words (real_type_inout).  `prev_last_row` is 0 for
 * partition 0, else last_row of the previous partition (:153-157).
 * `out` receives num_packets * 8 little-endian uint64 words.
 * The two out-of-range reads of coo_partition in the last packet (:208,:224)
 * are restated as "a different row".                                         */
ORC_API void orc_pack_partition(const uint32_t *row, const uint32_t *col, const uint32_t *val32,
                                uint64_t nnz_p, uint32_t prev_last_row, int W, uint64_t *out) {
    const int B = orc_packet_size(W);
    const uint64_t npk = (nnz_p + (uint64_t)B - 1) / (uint64_t)B;
    uint32_t curr_row = prev_last_row;
    for (uint64_t i = 0; i < npk; i++) {
        uint64_t *w = out + 8 * i;
        memset(w, 0, 64);
        uint64_t base = (uint64_t)B * i;
        uint32_t xl[32], yl[32], vl[32];
        /* xf: first nnz of the packet vs. last in-range nnz before it (:195-205) */
        uint32_t xf = (row[base] != curr_row) ? 1u : 0u;
        for (int j = 0; j < B; j++) {
            uint64_t g = base + (uint64_t)j;
            xl[j] = 0;
            if (g < nnz_p) {
                curr_row = row[g];
                yl[j] = col[g];
                vl[j] = orc_fxW_from_fx32(val32[g], W);
            } else {
                yl[j] = 0; vl[j] = 0;
            }
        }
        /* run lengths of equal rows -> x, then prefix-sum (:220-242) */
        int pos = 0;
        uint32_t same = 1;
        for (int j = 1; j < B; j++) {
            if (base + (uint64_t)j - 1 < nnz_p) {
                int same_row = (base + (uint64_t)j < nnz_p) && (row[base + j] == row[base + j - 1]);
                if (same_row) {
                    same++;
                } else {
                    xl[pos] = same; same = 1; pos++;
                }
            } else {
                xl[pos] = 0; pos++;
            }
        }
        if (base + (uint64_t)B - 1 < nnz_p) xl[pos] = same;
        for (int j = 1; j < B; j++) xl[j] += xl[j - 1];
        for (int j = 0; j < B; j++) {
            set_bits(w, 4u * (unsigned)j, 4, xl[j] & 0xFu);
            set_bits(w, 4u * (unsigned)B + 10u * (unsigned)j, 10, yl[j] & 0x3FFu);
            set_bits(w, 14u * (unsigned)B + (unsigned)W * (unsigned)j, (unsigned)W, vl[j]);
        }
        set_bits(w, 511, 1, xf);
    }
}

/* fpga_utils.hpp:346-355 write_block_vec + host_spmv_bscsr.cpp:173-186:
 * B 32-bit words per 64-byte block, zero padded to ceil(C/B) blocks.         */
ORC_API void orc_pack_query(const uint32_t *vec32, uint32_t cols, int W, uint64_t *out) {
    const int B = orc_packet_size(W);
    uint32_t nblk = (cols + (uint32_t)B - 1) / (uint32_t)B;
    for (uint32_t i = 0; i < nblk; i++) {
        uint64_t *w = out + 8 * (uint64_t)i;
        memset(w, 0, 64);
        for (int j = 0; j < B; j++) {
            uint32_t c = i * (uint32_t)B + (uint32_t)j;
            uint32_t v = (c < cols) ? vec32[c] : 0u;
            set_bits(w, 32u * (unsigned)j, 32, v);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* HLS kernel, one sub-core (= one partition), literal sequential form       */
/* ------------------------------------------------------------------------ */

/* spmv_bscsr_top_k_multicore.hpp:28-99 argmin_K with MIN(res,a,b) =
 * res[a] < res[b] ? a : b  (ties -> b, the higher slot).  K==4 keeps the
 * reference's `MIN(res, 2, 2)` typo (:45): slot 3 is never the minimum.      */
static int orc_argmin(const uint32_t *res, int K) {
    if (K == 4) {
        int m0 = (res[0] < res[1]) ? 0 : 1;
        int m1 = 2;
        return (res[m0] < res[m1]) ? m0 : m1;
    }
    if (K == 1) return 0;
    if (K == 2 || K == 8 || K == 16) {
        int idx[16];
        int n = K;
        for (int i = 0; i < n; i++) idx[i] = i;
        while (n > 1) {
            for (int i = 0; i < n / 2; i++) {
                int a = idx[2 * i], b = idx[2 * i + 1];
                idx[i] = (res[a] < res[b]) ? a : b;
            }
            n /= 2;
        }
        return idx[0];
    }
    int cm = 0;                                   /* template argmin<k> (:89-99) */
    for (int i = 0; i < K; i++) cm = (res[cm] < res[i]) ? cm : i;
    return cm;
}

/* spmv_bscsr_top_k_multicore.cpp:112-185 (vec load, write-back) and
 * spmv_bscsr_top_k_multicore.hpp:104-149, 168-220, 246-326, 331-409.
 *   packets : npk * 8 uint64 words of one partition
 *   xq32    : `cols` raw 32-bit query words (what write_block_vec stored)
 *   out_idx / out_val : Kp words of 16 x u32 each, reference result layout
 *             (lane q at position q; positions >= LFR stay 0; values widened
 *             to 32-bit fixed, i.e. << (32-W)).
 * Column indices >= MAX_COLS cannot occur (10-bit field); a column >= cols
 * reads an entry of the URAM copy that COPY_INPUT never wrote; the static
 * array is zero-initialised, so it reads 0.                                  */
/* drift_free != 0 is NOT the reference: it is the stated repair of SURVEY 7-H2 that the engine offers as
 * TKS fixed_drift_free.  Two rules change, both only visible in packets with more than LFR row segments:
 *   (1) the row counter advances by the TRUE number of rows that finish in the packet (all B segment-end
 *       fields), so reported row indices never drift;
 *   (2) the partial sum carried to the next packet is that of the packet's true last segment (the reference
 *       carries agg[LFR-1], the sum of a row that has already finished).
 * Rows that finish in such a packet without a lane (segments LFR-1 .. last-1) are still not offered.       */
ORC_API void orc_bscsr_partition_ex(const uint64_t *packets, uint64_t npk,
                                    const uint32_t *xq32, uint32_t cols,
                                    int W, int Kp, int LFR, int drift_free,
                                    uint32_t *out_idx, uint32_t *out_val) {
    const int B = orc_packet_size(W);
    const int F = W - 1;
    const uint64_t M = (W == 32) ? 0xFFFFFFFFull : ((1ull << W) - 1ull);
    uint32_t xq[1024];
    for (uint32_t c = 0; c < 1024; c++) xq[c] = (c < cols) ? (xq32[c] >> (32 - W)) : 0u;

    /* res_local / res_idx_local [LFR][Kp], curr_worst_* [LFR]  (:154-163, :490-500) */
    uint32_t Lval[16][64], Lidx[16][64];
    int worst_idx[16]; uint32_t worst_val[16];
    for (int j = 0; j < LFR; j++) {
        for (int t = 0; t < Kp; t++) { Lval[j][t] = 0; Lidx[j][t] = 0; }
        worst_idx[j] = 0; worst_val[j] = 0;
    }
    uint32_t last_row = 0;         /* last_row_of_packet (:260)        */
    uint32_t last_out = 0;         /* last_row_of_packet_output (:261) */

    for (uint64_t i = 0; i < npk; i++) {
        const uint64_t *w = packets + 8 * i;
        uint32_t x[32], pw[32];
        for (int j = 0; j < B; j++) {
            x[j] = (uint32_t)get_bits(w, 4u * (unsigned)j, 4);
            uint32_t y = (uint32_t)get_bits(w, 4u * (unsigned)B + 10u * (unsigned)j, 10);
            uint32_t v = (uint32_t)get_bits(w, 14u * (unsigned)B + (unsigned)W * (unsigned)j, (unsigned)W);
            /* ufixed<W,1> * ufixed<W,1> -> full precision, assigned to
             * ufixed<W,1,AP_TRN_ZERO>: drop F low bits, wrap to W (:121-126) */
            pw[j] = (uint32_t)((((uint64_t)v * (uint64_t)xq[y]) >> F) & M);
        }
        uint32_t xf = (uint32_t)get_bits(w, 511, 1);

        /* loop 2 (:128-146) */
        uint32_t agg[16]; uint32_t n = 0;
        for (int s = 0; s < LFR; s++) {
            uint32_t st = (s > 0) ? x[s - 1] : 0u;
            uint32_t en = x[s];
            n += (st != en);
            uint64_t a = 0;
            for (int j = 0; j < B; j++) if ((uint32_t)j >= st && (uint32_t)j < en) a = (a + pw[j]) & M;
            agg[s] = (uint32_t)a;
        }

        /* loop 3 (:268-308) */
        uint32_t al[17]; int fin[17];
        for (int j = 0; j <= LFR; j++) { al[j] = 0; fin[j] = 0; }
        uint32_t nw = (i != 0) ? xf : 0u;
        uint32_t finished_rows_num = n + nw - 1u;          /* int_type arithmetic, wraps */
        uint32_t nseg = 0, last_sum = 0;
        if (drift_free) {
            uint32_t st = 0, last_st = 0, last_en = 0;
            for (int s2 = 0; s2 < B; s2++) {
                if (x[s2] != st) { nseg++; last_st = st; last_en = x[s2]; }
                st = x[s2];
            }
            uint64_t a2 = 0;
            for (uint32_t j = last_st; j < last_en; j++) a2 = (a2 + pw[j]) & M;
            last_sum = (uint32_t)a2;
            if (nseg > (uint32_t)LFR) finished_rows_num = nseg + nw - 1u;
        }
        uint32_t start_row = last_row + nw;
        last_row += finished_rows_num;
        al[1] = agg[0];
        for (int j = 1; j < LFR; j++) {
            al[1 + j] = agg[j];
            fin[j] = (x[j - 1] != ((j > 1) ? x[j - 2] : 0u));
        }
        fin[n] = 0;
        if (!nw) {
            al[1] = (uint32_t)(((uint64_t)al[1] + last_out) & M);
            al[0] = 0; fin[0] = 0;
        } else {
            al[0] = last_out; fin[0] = 1;
        }
        last_out = al[n];
        if (drift_free && nseg > (uint32_t)LFR) last_out = last_sum;

        /* loop 4 (:366-389) */
        for (int j = 0; j < LFR; j++) {
            uint32_t cv = al[j];
            if (cv >= worst_val[j] && fin[j]) {
                Lidx[j][worst_idx[j]] = start_row + (uint32_t)j - 1u;
                Lval[j][worst_idx[j]] = cv;
            }
            worst_idx[j] = orc_argmin(Lval[j], Kp);
            worst_val[j] = Lval[j][worst_idx[j]];
        }
    }
    /* no flush of the last row (:392-399 is commented out) */

    /* write-back (.cpp:151-185): word t, position j = list j slot t */
    for (int t = 0; t < Kp; t++) {
        for (int q = 0; q < 16; q++) {
            uint32_t vi = 0, vv = 0;
            if (q < LFR && q < B) { vi = Lidx[q][t]; vv = (W == 32) ? Lval[q][t] : (Lval[q][t] << (32 - W)); }
            out_idx[16 * t + q] = vi;
            out_val[16 * t + q] = vv;
        }
    }
}

ORC_API void orc_bscsr_partition(const uint64_t *packets, uint64_t npk,
                                 const uint32_t *xq32, uint32_t cols,
                                 int W, int Kp, int LFR,
                                 uint32_t *out_idx, uint32_t *out_val) {
    orc_bscsr_partition_ex(packets, npk, xq32, cols, W, Kp, LFR, 0, out_idx, out_val);
}

/* host_spmv_bscsr.cpp:399-448 read_result + evaluation_utils.hpp:40-62.
 * Candidates of all partitions: idx += first_row[p]; keep val > 0; first
 * insertion of an index wins (unordered_map::insert), visiting order is
 * partition -> slot t -> position q; then sort (val desc, idx desc).
 * res_* must hold P*Kp*B entries.  Returns the number of results.           */
ORC_API uint32_t orc_read_result(int P, int Kp, int B, const uint32_t *idx_words,
                                 const uint32_t *val_words, const uint32_t *first_row,
                                 uint32_t *res_idx, uint32_t *res_val) {
    uint32_t cnt = 0;
    for (int p = 0; p < P; p++) {
        for (int t = 0; t < Kp; t++) {
            for (int q = 0; q < B; q++) {
                uint64_t o = ((uint64_t)p * (uint64_t)Kp + (uint64_t)t) * 16u + (uint64_t)q;
                uint32_t id = idx_words[o] + first_row[p];
                uint32_t v = val_words[o];
                if (v > 0) {
                    int dup = 0;
                    for (uint32_t e = 0; e < cnt; e++) if (res_idx[e] == id) { dup = 1; break; }
                    if (!dup) { res_idx[cnt] = id; res_val[cnt] = v; cnt++; }
                }
            }
        }
    }
    orc_sort_tuples_u32(cnt, res_idx, res_val);
    return cnt;
}

/* gold_algorithms.hpp:188-246 instantiated with V = ap_ufixed<32,1,AP_TRN_ZERO>
 * (the FPGA host's software reference, host_spmv_bscsr.cpp:487-505).         */
ORC_API void orc_gold_topk_fx32(const uint32_t *row, const uint32_t *col, const uint32_t *val32,
                                uint64_t nnz, const uint32_t *vec32, int k,
                                uint32_t *res_idx, uint32_t *res_val) {
    for (int i = 0; i < k; i++) { res_idx[i] = 0; res_val[i] = 0; }
    if (nnz == 0) return;
    uint32_t curr_row = row[0], curr_out = 0, worst_idx = 0, worst_val = 0;
    for (uint64_t i = 0; i < nnz; i++) {
        uint32_t r = row[i];
        uint32_t contrib = (uint32_t)(((uint64_t)val32[i] * (uint64_t)vec32[col[i]]) >> 31);
        if (r == curr_row) {
            curr_out += contrib;
        } else {
            if (curr_out >= worst_val) {
                res_idx[worst_idx] = curr_row;
                res_val[worst_idx] = curr_out;
                uint32_t w = 0, wv = res_val[0];
                for (int j = 0; j < k; j++) if (res_val[j] < wv) { w = (uint32_t)j; wv = res_val[j]; }
                worst_idx = w; worst_val = wv;
            }
            curr_row = r;
            curr_out = contrib;
        }
    }
    if (curr_out >= worst_val) { res_idx[worst_idx] = curr_row; res_val[worst_idx] = curr_out; }
}
