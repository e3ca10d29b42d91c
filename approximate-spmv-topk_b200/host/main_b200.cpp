// main_b200.cpp -- host executable `build/topk-spmv-b200`, the drop-in for the reference's per-backend
// hosts (src/fpga/src/host_spmv_bscsr.cpp:510-707, src/gpu/host_spmv_topk_csr_gpu.cu:291-480):
//   Options -> readMtx -> coo_t -> create_sample_vector -> software gold -> SpMV ctor
//   -> for each test: new query, gold, reset, operator(), read_result, error counters, one CSV row.
// Same CLI (-m matrix -k top-K -t iterations -d debug ...), same CSV columns in the same order, so that
// test_spmv_topk.py and the plotting scripts' readers keep working; throughput columns are appended at
// the END of each row.  The accelerator behind `SpMV` is libtopkspmv.so (B200, sm_100a).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <tuple>
#include <unordered_set>
#include <vector>

#include "coo_matrix.hpp"
#include "evaluation.hpp"
#include "fixed_point.hpp"
#include "gold.hpp"
#include "mtx_reader.hpp"
#include "options.hpp"
#include "spmv.hpp"
#include "types.hpp"

namespace chrono = std::chrono;
using clock_type = chrono::high_resolution_clock;

static float to_print(float v) { return v; }
static float to_print(ufixed32 v) { return v.to_float(); }

template <typename V>
static std::string join_vals(const std::vector<V> &v) {
    std::string s;
    for (size_t j = 0; j < v.size(); j++) s += std::to_string(to_print(v[j])) + (j + 1 < v.size() ? ";" : "");
    return s;
}
static std::string join_idx(const std::vector<int_type> &v) {
    std::string s;
    for (size_t j = 0; j < v.size(); j++) s += std::to_string(v[j]) + (j + 1 < v.size() ? ";" : "");
    return s;
}

// sw_test (host_spmv_bscsr.cpp:487-505): the reference's software result, timed
template <typename V>
static float sw_test(const coo_t<int_type, V> &coo, std::vector<int_type> &res_idx_sw, std::vector<V> &res_sim_sw, V *vec, int k) {
    auto t0 = clock_type::now();
    spmv_coo_gold_top_k<int_type, V>(coo, vec, k, res_idx_sw.data(), res_sim_sw.data());
    sort_tuples(k, res_idx_sw.data(), res_sim_sw.data());
    return (float)chrono::duration_cast<chrono::microseconds>(clock_type::now() - t0).count() / 1000;
}

template <typename V, typename Engine>
static int run(const Options &options, coo_t<int_type, V> &coo, int_type rows, int_type cols, int_type nnz, Engine &spmv,
               std::vector<V> &vec, float setup_ms, const char *time_cols, double algorithmic_bytes) {
    const int debug = options.debug;
    const int k = options.top_k_value;
    std::vector<V> res_sim_sw(k);
    std::vector<int_type> res_idx_sw(k);
    float sw_time = sw_test(coo, res_idx_sw, res_sim_sw, vec.data(), k);
    std::vector<float> exec_times, exec_times_full, readback_times, error_count, precision_vec;
    for (unsigned i = 0; i < options.num_tests; i++) {
        if (debug) std::cout << "\nIteration " << i << ")" << std::endl;
        if (options.reset) {
            create_sample_vector(vec.data(), (int)cols, true, false, true, options.seed ? options.seed + (int)i + 1 : 0);
            sw_time = sw_test(coo, res_idx_sw, res_sim_sw, vec.data(), k);
            spmv.reset(vec.data(), debug);
        }
        std::vector<V> hw_res;
        std::vector<int_type> hw_res_idx;
        auto t5 = clock_type::now();
        float exec_ms = (float)spmv(debug) / 1e6f;
        float full_ms = (float)chrono::duration_cast<chrono::nanoseconds>(clock_type::now() - t5).count() / 1e6f;
        exec_times.push_back(exec_ms);
        exec_times_full.push_back(full_ms);
        auto t6 = clock_type::now();
        spmv.read_result(hw_res, hw_res_idx, debug);
        float readback_ms = (float)chrono::duration_cast<chrono::microseconds>(clock_type::now() - t6).count() / 1000;
        readback_times.push_back(readback_ms);

        // correctness counters exactly as host_spmv_bscsr.cpp:636-650
        const int res_size = (int)hw_res_idx.size();
        int error_idx = check_array_equality(hw_res_idx.data(), res_idx_sw.data(), std::min(k, res_size), 0, debug);
        error_count.push_back((float)error_idx);
        error_idx += std::max(0, k - res_size);
        int error = check_array_equality(hw_res.data(), res_sim_sw.data(), std::min(k, res_size), 10e-6, debug);
        error += std::max(0, k - res_size);
        std::unordered_set<int_type> s(res_idx_sw.begin(), res_idx_sw.end());
        int inter = (int)std::count_if(hw_res_idx.begin(), hw_res_idx.end(), [&](int_type r) { return s.count(r) != 0; });
        float precision = (float)inter / (float)k;
        precision_vec.push_back(precision);
        const double nnz_per_s = exec_ms > 0 ? (double)nnz / (exec_ms * 1e-3) : 0.0;
        const double gbs = exec_ms > 0 ? algorithmic_bytes / (exec_ms * 1e-3) / 1e9 : 0.0;
        if (debug) {
            std::cout << "sw results =" << std::endl;
            for (int j = 0; j < k; j++) std::cout << j << ") document " << res_idx_sw[j] << " = " << res_sim_sw[j] << std::endl;
            std::cout << "hw results=" << std::endl;
            for (int j = 0; j < std::min(k, res_size); j++) std::cout << j << ") document " << hw_res_idx[j] << " = " << hw_res[j] << std::endl;
            std::cout << "num errors on indices=" << error_idx << "\nnum errors on values=" << error << "\nprecision=" << precision << std::endl;
            std::cout << "b200 exec time=" << exec_ms << " ms, full exec time=" << full_ms << " ms, " << nnz_per_s / 1e9 << " Gnnz/s, " << gbs << " GB/s" << std::endl;
        } else {
            if (i == 0)
                std::cout << "iteration,error_idx,error_val,sw_full_time_ms,sw_topk_time_ms,hw_setup_time_ms," << time_cols
                          << ",readback_time_ms,k,sw_res_idx,sw_res_val,hw_res_idx,hw_res_val,precision,nnz_per_s,effective_gb_per_s" << std::endl;
            std::cout << i << "," << error_idx << "," << error << "," << 0 << "," << sw_time << "," << setup_ms << "," << exec_ms << "," << full_ms
                      << "," << readback_ms << "," << k << "," << join_idx(res_idx_sw) << "," << join_vals(res_sim_sw) << ","
                      << join_idx(hw_res_idx) << "," << join_vals(hw_res) << "," << precision << "," << nnz_per_s << "," << gbs << std::endl;
        }
    }
    if (debug) {
        int old = (int)std::cout.precision();
        std::cout.precision(4);
        std::cout << "----------------" << std::endl;
        std::cout << "Mean B200 execution time=" << mean(exec_times, 2) << "±" << st_dev(exec_times, 2) << " ms" << std::endl;
        std::cout << "Mean full B200 execution time=" << mean(exec_times_full, 2) << "±" << st_dev(exec_times_full, 2) << " ms" << std::endl;
        std::cout << "Mean read-back time=" << mean(readback_times, 2) << "±" << st_dev(readback_times, 2) << " ms" << std::endl;
        std::cout << "Mean error=" << mean(error_count, 2) << "±" << st_dev(error_count, 2) << std::endl;
        std::cout << "Mean precision=" << mean(precision_vec, 2) << "±" << st_dev(precision_vec, 2) << std::endl;
        std::cout << "----------------" << std::endl;
        std::cout.precision(old);
    }
    if (options.throughput_queries > 0) {
        // -F n: the loop above keeps ONE query in flight, like the reference's hosts.  Here n fresh queries go through
        // the throughput verbs (submit / fetch, Engine::kInFlight queries behind); every result must equal what the
        // blocking verbs return for the same query.  Reported on stderr so that the CSV on stdout keeps its schema.
        const unsigned n = options.throughput_queries;
        std::vector<std::vector<V>> qs(n, std::vector<V>(cols));
        std::vector<std::vector<int_type>> want_idx(n);
        std::vector<std::vector<V>> want_val(n);
        for (unsigned i = 0; i < n; i++) {
            create_sample_vector(qs[i].data(), (int)cols, true, false, true, options.seed ? options.seed + 1000 + (int)i : 0);
            spmv.reset(qs[i].data(), 0);
            spmv(0);
            spmv.read_result(want_val[i], want_idx[i], 0);
        }
        std::vector<uint64_t> tickets(n);
        std::vector<V> got_val;
        std::vector<int_type> got_idx;
        unsigned mismatches = 0;
        const unsigned lag = (unsigned)Engine::kInFlight;
        auto check = [&](unsigned i) {
            spmv.fetch(tickets[i], got_val, got_idx);
            bool same = got_idx.size() == want_idx[i].size();
            for (size_t j = 0; j < got_idx.size() && same; j++) same = got_idx[j] == want_idx[i][j] && to_print(got_val[j]) == to_print(want_val[i][j]);
            mismatches += same ? 0u : 1u;
        };
        auto t0 = clock_type::now();
        for (unsigned i = 0; i < n; i++) {
            tickets[i] = spmv.submit(qs[i].data());
            if (i >= lag) check(i - lag);
        }
        for (unsigned i = (n > lag ? n - lag : 0); i < n; i++) check(i);
        const double ms = (double)chrono::duration_cast<chrono::nanoseconds>(clock_type::now() - t0).count() / 1e6 / n;
        std::cerr << "throughput: " << n << " queries through submit/fetch, " << ms << " ms per query, "
                  << (double)nnz / (ms * 1e-3) / 1e9 << " Gnnz/s, results differing from the blocking verbs: " << mismatches << std::endl;
        if (mismatches) return 2;
    }
    (void)rows;
    return 0;
}

int main(int argc, char *argv[]) {
    Options options(argc, argv);
    const int debug = options.debug;
    int_type nnz = 0, rows = 0, cols = 0;
    std::vector<int_type> x, y;
    std::vector<double> val_d;
    auto t1 = clock_type::now();
    std::string err;
    const char *path = options.use_sample_matrix ? DEFAULT_MTX_FILE : options.matrix_path.c_str();
    // -C: the parsed matrix is kept as a checksummed binary CSR next to the text file (tks_cache_*); a sweep re-reads
    // that instead of the Matrix-Market text
    // The cache holds fp32 values, so only the float engine reads it (the fixed-point engine quantises the parsed
    // doubles, utils.hpp:401, and must not depend on whether an earlier float run left a cache behind); and it is used
    // only if it was made from THIS file as it is now, with the same -z / -v flags.
    bool from_cache = false;
    const uint64_t want_tag = tks_cache_source_tag(path, options.zero_indexed, options.ignore_matrix_values);
    if (!options.cache_path.empty() && options.use_float) {
        uint64_t crow = 0, cnnz = 0, have_tag = 0;
        uint32_t ccol = 0;
        const int rc_size = tks_cache_read_csr_tagged(options.cache_path.c_str(), &crow, &ccol, &cnnz, nullptr, nullptr, nullptr, &have_tag);
        if (rc_size == TKS_OK && have_tag != want_tag && debug)
            std::cout << "matrix cache " << options.cache_path << " was made from another file, an older version of it or other -z/-v flags: ignored" << std::endl;
        if (rc_size == TKS_OK && have_tag == want_tag) {
            std::vector<uint64_t> p64(crow + 1);
            std::vector<float> v32(cnnz);
            y.resize(cnnz);
            if (tks_cache_read_csr(options.cache_path.c_str(), &crow, &ccol, &cnnz, p64.data(), y.data(), v32.data()) != TKS_OK) {
                std::cerr << "matrix cache " << options.cache_path << " is unusable" << std::endl;
                return 1;
            }
            rows = (int_type)crow; cols = ccol; nnz = (int_type)cnnz;
            x.resize(cnnz);
            for (uint64_t r = 0; r < crow; r++)
                for (uint64_t e = p64[r]; e < p64[r + 1]; e++) x[e] = (int_type)r;
            val_d.assign(v32.begin(), v32.end());
            from_cache = true;
        }
    }
    if (!from_cache) {
        if (tkshost::readMtx<int_type, double>(path, &x, &y, &val_d, &rows, &cols, &nnz, 0, !options.ignore_matrix_values, debug,
                                               options.zero_indexed, false, &err) != 0) {
            std::cerr << err << std::endl;
            return 1;
        }
        if (!options.cache_path.empty() && options.use_float) {
            // the cache holds what the float engine consumes: CSR with fp32 values
            std::vector<float> v32(val_d.begin(), val_d.end());
            std::vector<int_type> ptr32(rows + 1), idx(nnz);
            std::vector<float> cv(nnz);
            if (tkshost::coo2csr<int_type, float>(ptr32.data(), idx.data(), cv.data(), x, y, v32, rows, cols) == 0) {
                std::vector<uint64_t> p64(ptr32.begin(), ptr32.end());
                if (tks_cache_write_csr_tagged(options.cache_path.c_str(), rows, cols, nnz, p64.data(), idx.data(), cv.data(), want_tag) != TKS_OK)
                    std::cerr << "warning: could not write matrix cache " << options.cache_path << std::endl;
            }
        }
    }
    if (debug) {
        auto ms = chrono::duration_cast<chrono::milliseconds>(clock_type::now() - t1).count();
        std::cout << "loaded matrix with " << rows << " rows, " << cols << " columns and " << nnz << " non-zero elements" << std::endl;
        std::cout << "setup time=" << ms << " ms" << std::endl;
    }

    if (options.use_float) {
        std::vector<float> val(val_d.begin(), val_d.end());
        coo_t<int_type, float> coo(x, y, val);
        std::vector<int_type> ptr(rows + 1), idx(nnz);
        std::vector<float> csr_val(nnz);
        if (tkshost::coo2csr<int_type, float>(ptr.data(), idx.data(), csr_val.data(), x, y, val, rows, cols) != 0) {
            std::cerr << "Error: Index out of bounds!" << std::endl;
            return 1;
        }
        std::vector<float> vec(cols);
        create_sample_vector(vec.data(), (int)cols, true, false, true, options.seed);
        auto t4 = clock_type::now();
        const double bytes_g = (options.use_half_precision_gpu ? 6.0 : 8.0) * nnz + 4.0 * (rows + 1.0) + 4.0 * cols + 8.0 * options.top_k_value;
        if (options.devices.size() > 1) {
            // -G a,b,...: one shard per listed device, all driven by this process (tks_group_*); same loop, same CSV
            SpMVGroup group(ptr.data(), idx.data(), csr_val.data(), rows, cols, nnz, vec.data(), options.top_k_value,
                            options.devices, options.tie_higher, debug, options.use_half_precision_gpu);
            float setup_g = (float)chrono::duration_cast<chrono::microseconds>(clock_type::now() - t4).count() / 1000;
            if (debug) std::cout << "b200 setup time=" << setup_g << " ms (" << options.devices.size() << " devices)" << std::endl;
            return run<float>(options, coo, rows, cols, nnz, group, vec, setup_g, "hw_spmv_only_time_ms,hw_exec_time_ms", bytes_g);
        }
        SpMV spmv(ptr.data(), idx.data(), csr_val.data(), rows, cols, nnz, vec.data(), options.top_k_value, options.device,
                  options.tie_higher, debug, options.use_half_precision_gpu);   // -a, as host_spmv_topk_csr_gpu.cu:382
        float setup_ms = (float)chrono::duration_cast<chrono::microseconds>(clock_type::now() - t4).count() / 1000;
        if (debug) std::cout << "b200 setup time=" << setup_ms << " ms" << std::endl;
        const double bytes = (options.use_half_precision_gpu ? 6.0 : 8.0) * nnz + 4.0 * (rows + 1.0) + 4.0 * cols + 8.0 * options.top_k_value;
        // GPU-host column names (host_spmv_topk_csr_gpu.cu:452)
        return run<float>(options, coo, rows, cols, nnz, spmv, vec, setup_ms, "hw_spmv_only_time_ms,hw_exec_time_ms", bytes);
    }
    std::vector<ufixed32> val(val_d.begin(), val_d.end());   // (T) value, utils.hpp:401
    coo_t<int_type, ufixed32> coo(x, y, val);
    std::vector<ufixed32> vec(cols);
    create_sample_vector(vec.data(), (int)cols, true, false, true, options.seed);
    auto t4 = clock_type::now();
    SpMVFixed spmv(x.data(), y.data(), val.data(), rows, cols, nnz, vec.data(), options.top_k_value, options.fixed_width,
                   options.partitions, options.local_k, options.limited_finished_rows, options.device, debug,
                   options.drift_free, options.device_pack);
    float setup_ms = (float)chrono::duration_cast<chrono::microseconds>(clock_type::now() - t4).count() / 1000;
    if (debug) std::cout << "b200 setup time=" << setup_ms << " ms" << std::endl;
    tks_stats st;
    tks_get_stats(spmv.h, &st);
    // FPGA-host column names (host_spmv_bscsr.cpp:667)
    return run<ufixed32>(options, coo, rows, cols, nnz, spmv, vec, setup_ms, "hw_exec_time_ms,hw_full_exec_time_ms",
                         (double)st.algorithmic_bytes);
}
