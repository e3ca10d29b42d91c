"""Row sharding for multi-GPU runs.

Rows are independent and top-k(union) = top-k(union of per-shard top-k), so the matrix is split into
contiguous row ranges, one per rank, exactly like the reference splits rows over its 32 FPGA
partitions (src/fpga/src/host_spmv_bscsr.cpp:136-141) -- except that boundaries are balanced by
non-zeros, not by row count.  Each rank reports GLOBAL row ids (row_offset of tks_upload_csr), the K
candidates per rank are all-gathered (K * 8 bytes) and merged on every rank.
"""
from __future__ import annotations

import numpy as np


def plan_row_shards_by_nnz(ptr, n_shards):
    """ptr: CSR row pointer (rows+1).  Returns [(row_begin, row_end)] with ~equal nnz, contiguous, covering."""
    ptr = np.asarray(ptr, dtype=np.uint64)
    rows = ptr.size - 1
    nnz = int(ptr[-1])
    bounds = [0]
    for s in range(1, n_shards):
        target = nnz * s // n_shards
        r = int(np.searchsorted(ptr, target, side="left"))
        r = min(max(r, bounds[-1]), rows)
        bounds.append(r)
    bounds.append(rows)
    return [(bounds[i], bounds[i + 1]) for i in range(n_shards)]


def plan_row_shards_even(rows, n_shards):
    """Equal row counts (used for matrices generated on the device, whose degrees are i.i.d.)."""
    base, rem = divmod(int(rows), n_shards)
    out, b = [], 0
    for s in range(n_shards):
        e = b + base + (1 if s < rem else 0)
        out.append((b, e))
        b = e
    return out


def slice_csr(ptr, idx, val, row_begin, row_end):
    ptr = np.asarray(ptr)
    b, e = int(ptr[row_begin]), int(ptr[row_end])
    return (ptr[row_begin:row_end + 1] - ptr[row_begin]).astype(np.uint64), idx[b:e], val[b:e]


def merge_topk_host(keys_per_rank, k):
    """Reference merge on the host (used by the gloo CPU tests of the plumbing): keys are the engine's
    64-bit ordering keys (score bits << 32 | tie-ordered index); the k largest, descending."""
    allk = np.concatenate([np.asarray(a, np.uint64) for a in keys_per_rank])
    allk = allk[allk != 0]
    return np.sort(allk)[::-1][:k]


def make_keys(scores, rows, tie_higher=False):
    """Host replica of common.cuh make_key for float32 scores (plumbing tests only)."""
    b = np.asarray(scores, np.float32).view(np.uint32).astype(np.uint64)
    neg = (b >> np.uint64(31)) != 0
    ordered = np.where(neg, (~b) & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))
    r = np.asarray(rows, np.uint32).astype(np.uint64)
    lo = r if tie_higher else ((~r) & np.uint64(0xFFFFFFFF))
    return (ordered << np.uint64(32)) | lo


def split_keys(keys, tie_higher=False):
    keys = np.asarray(keys, np.uint64)
    ordered = (keys >> np.uint64(32)).astype(np.uint32)
    lo = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    rows = lo if tie_higher else ~lo
    neg = (ordered & np.uint32(0x80000000)) == 0
    bits = np.where(neg, ~ordered, ordered & np.uint32(0x7FFFFFFF)).astype(np.uint32)
    return bits.view(np.float32), rows
