// options.hpp -- command line of the host executable.  Same struct name, field names, defaults and
// flags as the reference (src/common/utils/options.hpp:37-132: -d -s -r -m -t -x -v -k -b -c -g -i -a), so
// test_spmv_topk.py's command templates keep working; flags that only made sense for the FPGA/cuSPARSE
// back-ends (-x -b -c -g -i) are accepted and ignored; -a selects half-precision values as in the reference.  New flags select what the reference fixed at
// compile time (types.hpp) or hard-coded in main():
//   -z  the MTX file is 0-indexed          (reference hosts assume this, host_spmv_bscsr.cpp:539)
//   -f  FPGA-semantics fixed-point BS-CSR engine instead of exact fp32 CSR (USE_FLOAT, types.hpp:29)
//   -w  FIXED_WIDTH   -p SPMV_PARTITIONS   -l LIMITED_FINISHED_ROWS   -q local K   (types.hpp:20,36,77,51)
//   -e  seed of the query generator (0 = random_device, as the reference)
//   -G  CUDA device ordinal, or a list "0,1,...,7": rows sharded over those GPUs, driven by this one process (float engine)
//   -T  ties -> higher index first (reference sort order)
//   -D  fixed-point engine: repair the reference's row-counter drift (tks_config.fixed_drift_free)
//   -P  fixed-point engine: build the BS-CSR packets on the GPU (tks_upload_coo_fixed) instead of on the host
//   -C  <file>  binary matrix cache: loaded when present, else written after the MTX text is parsed
#pragma once

#include <getopt.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "types.hpp"

#define DEBUG false
#define RESET true
#define DEFAULT_MTX_FILE "../../data/matrices_for_testing/matrices_small/matrix_1000_512_20_gamma.mtx"
#define DEFAULT_BLOCK_SIZE_1D 32
#define DEFAULT_BLOCK_SIZE_2D 8
#define DEFAULT_NUM_BLOCKS 64
#define DEFAULT_GPU_IMPL 0
#define DEFAULT_USE_HALF_PRECISION_GPU false
#define DEFAULT_NUM_TESTS 3
#define DEFAULT_TOP_K 20
#define XCLBIN "../approximate_spmv.xclbin"

enum GPU_IMPL { CSR = 0, CSR_LIGHTSPMV = 1, COO = 2 };

struct Options {
    // Input-specific options;
    std::string matrix_path = DEFAULT_MTX_FILE;
    bool use_sample_matrix = false;
    bool reset = RESET;
    // Testing options;
    unsigned num_tests = DEFAULT_NUM_TESTS;
    int debug = DEBUG;
    bool ignore_matrix_values = false;
    int top_k_value = DEFAULT_TOP_K;
    // Accepted for compatibility, unused by this engine;
    std::string xclbin_path = XCLBIN;
    GPU_IMPL gpu_impl = GPU_IMPL(DEFAULT_GPU_IMPL);
    bool use_half_precision_gpu = DEFAULT_USE_HALF_PRECISION_GPU;
    int block_size_1d = DEFAULT_BLOCK_SIZE_1D;
    int block_size_2d = DEFAULT_BLOCK_SIZE_2D;
    int num_blocks = DEFAULT_NUM_BLOCKS;
    // Runtime forms of the types.hpp knobs;
    bool zero_indexed = false;
    bool use_float = USE_FLOAT;
    int fixed_width = FIXED_WIDTH;
    int partitions = SPMV_PARTITIONS;
    int limited_finished_rows = LIMITED_FINISHED_ROWS;
    int local_k = K;
    int seed = 0;
    int device = 0;
    std::vector<int> devices;   // -G a,b,...: more than one entry = one shard per listed device (a device may repeat)
    bool tie_higher = false;
    bool drift_free = false;
    bool device_pack = false;
    std::string cache_path;
    unsigned throughput_queries = 0;   // -F n: after the tests, n more queries through submit / fetch (queries in flight)

    Options(int argc, char *argv[]) {
        static struct option long_options[] = {{"debug", no_argument, 0, 'd'},
                                               {"use_sample_matrix", no_argument, 0, 's'},
                                               {"no_reset", no_argument, 0, 'r'},
                                               {"matrix_path", required_argument, 0, 'm'},
                                               {"num_tests", required_argument, 0, 't'},
                                               {"xclbin", required_argument, 0, 'x'},
                                               {"ignore_matrix_values", no_argument, 0, 'v'},
                                               {"k", required_argument, 0, 'k'},
                                               {"block_size_1d", required_argument, 0, 'b'},
                                               {"block_size_2d", required_argument, 0, 'c'},
                                               {"num_blocks", required_argument, 0, 'g'},
                                               {"gpu_impl", required_argument, 0, 'i'},
                                               {"half_precision_gpu", no_argument, 0, 'a'},
                                               {"zero_indexed", no_argument, 0, 'z'},
                                               {"fixed", no_argument, 0, 'f'},
                                               {"fixed_width", required_argument, 0, 'w'},
                                               {"partitions", required_argument, 0, 'p'},
                                               {"limited_finished_rows", required_argument, 0, 'l'},
                                               {"local_k", required_argument, 0, 'q'},
                                               {"seed", required_argument, 0, 'e'},
                                               {"gpu", required_argument, 0, 'G'},
                                               {"tie_higher", no_argument, 0, 'T'},
                                               {"drift_free", no_argument, 0, 'D'},
                                               {"device_pack", no_argument, 0, 'P'},
                                               {"cache", required_argument, 0, 'C'},
                                               {"in_flight", required_argument, 0, 'F'},
                                               {0, 0, 0, 0}};
        int option_index = 0, opt;
        while ((opt = getopt_long(argc, argv, "dm:st:x:vk:rb:c:g:i:azfw:p:l:q:e:G:TDPC:F:", long_options, &option_index)) != EOF) {
            switch (opt) {
                case 'd': debug = true; break;
                case 'r': reset = true; break;   // sic: the reference's -r also sets true (options.hpp:90-92)
                case 'm': matrix_path = optarg; break;
                case 's': use_sample_matrix = true; break;
                case 't': num_tests = (unsigned)atoi(optarg); break;
                case 'x': xclbin_path = optarg; break;
                case 'v': ignore_matrix_values = true; break;
                case 'k': top_k_value = atoi(optarg); break;
                case 'b': block_size_1d = atoi(optarg); break;
                case 'c': block_size_2d = atoi(optarg); break;
                case 'g': num_blocks = atoi(optarg); break;
                case 'i': gpu_impl = GPU_IMPL(atoi(optarg)); break;
                case 'a': use_half_precision_gpu = true; break;
                case 'z': zero_indexed = true; break;
                case 'f': use_float = false; break;
                case 'w': fixed_width = atoi(optarg); break;
                case 'p': partitions = atoi(optarg); break;
                case 'l': limited_finished_rows = atoi(optarg); break;
                case 'q': local_k = atoi(optarg); break;
                case 'e': seed = atoi(optarg); break;
                case 'G': {
                    devices.clear();
                    for (const char *p = optarg; *p;) {
                        char *end = nullptr;
                        const long d = strtol(p, &end, 10);
                        if (end == p) break;
                        devices.push_back((int)d);
                        p = (*end == ',') ? end + 1 : end;
                        if (*end != ',' && *end != 0) break;
                    }
                    device = devices.empty() ? 0 : devices[0];
                    break;
                }
                case 'T': tie_higher = true; break;
                case 'D': drift_free = true; break;
                case 'C': cache_path = optarg; break;
                case 'P': device_pack = true; break;
                case 'F': throughput_queries = (unsigned)atoi(optarg); break;
                default: break;
            }
        }
    }
};
