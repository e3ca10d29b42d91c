// types.hpp -- the reference's compile-time knob surface (src/common/types.hpp), kept by name.
//
// In the reference these macros select one FPGA bitstream per build.  Here they are only the
// DEFAULTS of the host executable: every one of them is also a runtime field of tks_config
// (include/topkspmv.h), so one binary covers all designs of test_spmv_topk.py:40-47
// (32/26/21/20-bit fixed, float).  Override at build time with -DFIXED_WIDTH=32 etc.
#pragma once

#include <cstdint>

typedef unsigned int int_type;   // src/common/types.hpp:17
typedef int_type index_type;     // src/common/csc_matrix/csc_matrix.hpp:13

// Fixed-point format of the matrix values inside BS-CSR packets (types.hpp:20-22).
#ifndef FIXED_WIDTH
#define FIXED_WIDTH 20
#endif
#define SCALE (FIXED_WIDTH - 1)
#define FIXED_INTEGER_PART (FIXED_WIDTH - SCALE)

// Fixed-point format of queries / results on the host side (types.hpp:25-27).
#define FIXED_WIDTH_OUT 32
#define SCALE_OUT (FIXED_WIDTH_OUT - 1)
#define FIXED_INTEGER_PART_OUT (FIXED_WIDTH_OUT - SCALE_OUT)

// true = exact fp32 CSR engine, false = FPGA-semantics fixed-point BS-CSR engine (types.hpp:29).
#ifndef USE_FLOAT
#define USE_FLOAT true
#endif

#ifndef SPMV_PARTITIONS
#define SPMV_PARTITIONS 32        // types.hpp:36
#endif
#define SUB_SPMV_PARTITIONS 4     // types.hpp:37 (4 sub-cores per compute unit; layout only)
#define SUPER_SPMV_PARTITIONS (SPMV_PARTITIONS / SUB_SPMV_PARTITIONS)

#ifndef K
#define K 8                       // types.hpp:51 per-lane local top-K of the approximate mode
#endif
#define TOPK_RES_COPIES 1         // types.hpp:53
#ifndef MAX_COLS
#define MAX_COLS 1024             // types.hpp:55
#endif

#define AP_INT_VAL_BITWIDTH FIXED_WIDTH   // types.hpp:61-63
#define AP_INT_COL_BITWIDTH 10
#define AP_INT_ROW_BITWIDTH 4
#define DIM_BOOL 1
#define PACKET_TRIPLET_WIDTH (AP_INT_VAL_BITWIDTH + AP_INT_COL_BITWIDTH + AP_INT_ROW_BITWIDTH)
#define BSCSR_PORT_BITWIDTH 512           // types.hpp:71-73
#define BSCSR_PACKET_SIZE ((BSCSR_PORT_BITWIDTH - DIM_BOOL) / (PACKET_TRIPLET_WIDTH))
#define PADDING_SIZE (BSCSR_PORT_BITWIDTH - BSCSR_PACKET_SIZE * PACKET_TRIPLET_WIDTH + DIM_BOOL)

#ifndef LIMITED_FINISHED_ROWS
#define LIMITED_FINISHED_ROWS 4   // types.hpp:77 (set to BSCSR_PACKET_SIZE for the "clean" design, :76)
#endif
