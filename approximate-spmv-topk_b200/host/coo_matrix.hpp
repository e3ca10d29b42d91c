// coo_matrix.hpp -- host matrix container of the reference hosts (src/fpga/src/ip/coo_matrix.hpp:11-27):
// three parallel arrays plus num_rows = max(start) + 1 and num_nnz.
#pragma once

#include <algorithm>
#include <vector>

template <typename I, typename T>
struct coo_t {
    std::vector<I> start;   // row of every non-zero (row-sorted by construction of the inputs)
    std::vector<I> end;     // column
    std::vector<T> val;
    I num_rows = 0;
    I num_nnz = 0;

    coo_t(std::vector<I> s, std::vector<I> e, std::vector<T> v) : start(std::move(s)), end(std::move(e)), val(std::move(v)) {
        num_nnz = (I)start.size();
        I mx = 0;
        for (I r : start) mx = std::max(mx, r);
        num_rows = mx + 1;   // coo_matrix.hpp:23-26 (also 1 for an empty matrix)
    }
};
