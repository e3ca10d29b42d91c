// evaluation.hpp -- the few helpers of src/common/utils/evaluation_utils.hpp and utils.hpp that the
// Top-K hosts actually use: sort_tuples (:40-62), mean / st_dev with skip (:273-297),
// check_array_equality (utils.hpp:204-217), create_sample_vector (utils.hpp:234-267).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <numeric>
#include <random>
#include <vector>

// value descending; equal values -> HIGHER index first (the reference's order, evaluation_utils.hpp:52-56)
template <typename I, typename V>
inline void sort_tuples(size_t DIM, I *idx, V *val) {
    std::vector<size_t> perm(DIM);
    std::iota(perm.begin(), perm.end(), (size_t)0);
    std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) {
        if (val[a] != val[b]) return val[a] > val[b];
        return idx[a] > idx[b];
    });
    std::vector<I> i2(DIM);
    std::vector<V> v2(DIM);
    for (size_t i = 0; i < DIM; i++) { i2[i] = idx[perm[i]]; v2[i] = val[perm[i]]; }
    std::copy(i2.begin(), i2.end(), idx);
    std::copy(v2.begin(), v2.end(), val);
}

template <typename T>
inline T mean(const std::vector<T> &x, int skip = 0) {
    int n = (int)x.size() - skip;
    if (n <= 0) return (T)0;
    T sum = 0;
    for (size_t i = (size_t)skip; i < x.size(); i++) sum += x[i];
    return sum / (T)n;
}

template <typename T>
inline T st_dev(const std::vector<T> &x, int skip = 0) {
    int n = (int)x.size() - skip;
    if (n <= 0) return (T)0;
    T m = 0, m2 = 0;
    for (size_t i = (size_t)skip; i < x.size(); i++) { m += x[i]; m2 += x[i] * x[i]; }
    T diff = m2 - m * m / (T)n;
    if (diff < 0) diff = 0;
    return std::sqrt(diff / (T)n);
}

template <typename T>
inline int check_array_equality(T *x, T *y, int n, float tol = 0.0000001f, bool debug = false, int max_print = 20) {
    int num_errors = 0;
    for (int i = 0; i < n; i++) {
        float diff = (float)((x[i] > y[i]) ? (x[i] - y[i]) : (y[i] - x[i]));
        if (diff > tol) {
            num_errors++;
            if (debug && num_errors < max_print) std::cout << i << ") X: " << x[i] << ", Y: " << y[i] << ", diff: " << diff << std::endl;
        }
    }
    return num_errors;
}

// utils.hpp:234-267.  seed == 0 -> std::random_device, as in the reference.
template <typename T>
inline void create_sample_vector(T *vector, int size, bool random = false, bool sum_to_one = true, bool norm_one = false, int seed = 0) {
    if (random) {
        std::random_device rd;
        std::mt19937 engine(seed == 0 ? rd() : (unsigned)seed);
        std::uniform_real_distribution<double> dist(0, 1);
        for (int i = 0; i < size; i++) vector[i] = (T)dist(engine);
    } else {
        for (int i = 0; i < size; i++) vector[i] = (T)1;
    }
    if (sum_to_one) {
        float sum = 0;
        for (int i = 0; i < size; i++) sum += (float)vector[i];
        for (int i = 0; i < size; i++) vector[i] = (T)((float)vector[i] / sum);
    } else if (norm_one) {
        double sum = 0;
        for (int i = 0; i < size; i++) sum += (float)vector[i] * (float)vector[i];
        for (int i = 0; i < size; i++) vector[i] = (T)((float)vector[i] / std::sqrt(sum));
    }
}
