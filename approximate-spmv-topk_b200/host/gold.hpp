// gold.hpp -- the host executable's software self-check, as every reference host has one
// (sw_test: src/fpga/src/host_spmv_bscsr.cpp:487-505, src/gpu/host_spmv_topk_csr_gpu.cu:268-285).
// It fills the CSV columns sw_topk_time_ms / error_idx / error_val / sw_res_*; it is NOT a compute
// path of the engine (the engine has no CPU fallback) and it is independent of oracle/.
// Semantics of src/fpga/src/gold_algorithms/gold_algorithms.hpp:188-246: stream the row-sorted COO,
// keep k slots, a finished row takes over the worst slot when it is >= the worst value.
#pragma once

#include <vector>

template <typename I, typename V, typename COO>
inline void spmv_coo_gold_top_k(const COO &coo, const V *vec, int k, I *res_idx, V *res_val) {
    for (int i = 0; i < k; i++) { res_idx[i] = 0; res_val[i] = (V)0.0; }
    const size_t nnz = coo.start.size();
    if (nnz == 0) return;
    int worst = 0;
    V worst_val = (V)0.0;
    auto offer = [&](I row, V score, bool rescan) {
        if (!(score >= worst_val)) return;
        res_idx[worst] = row;
        res_val[worst] = score;
        if (!rescan) return;
        worst = 0;
        for (int j = 1; j < k; j++) if (res_val[j] < res_val[worst]) worst = j;   // first minimum
        worst_val = res_val[worst];
    };
    I row = coo.start[0];
    V acc = (V)0.0;
    for (size_t i = 0; i < nnz; i++) {
        const V contrib = coo.val[i] * vec[coo.end[i]];
        if (coo.start[i] == row) {
            acc += contrib;
        } else {
            offer(row, acc, true);
            row = coo.start[i];
            acc = contrib;
        }
    }
    offer(row, acc, false);   // the reference does not rescan after the final row (:239-244)
}
