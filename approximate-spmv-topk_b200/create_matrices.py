"""Synthetic sparse-embedding matrices with the law of the reference generator
(src/resources/python/create_matrices.py:57-130), restated with seeded, vectorised NumPy
(the reference needs `ray` and is unseeded):

  * row degrees   uniform: randint(avg//2, int(1.5*avg)+1)            (:84-86)
                  gamma  : max(int(Gamma(shape=3, scale=avg/3)), 1)     (:31, :91)
  * columns       sorted randint(0, max_cols) WITH replacement per row  (:45)
  * values        U[0,1), each row divided by its L2 norm               (:49, :54, :104)
  * MTX output    1-indexed, 10 significant digits, the exact 3-line header (:33, :120, :124)
                  so that test_cpu.py's `lines[2]` size parse keeps working.

    python approximate-spmv-topk_b200/create_matrices.py -r 10000 -c 1024 -n 20 -d gamma -o out_dir
"""
from __future__ import annotations

import argparse
import os

import numpy as np

GAMMA_K = 3
DEFAULT_PRECISION = 10
MTX_HEADER = "%%MatrixMarket matrix coordinate real general\n%\n{} {} {}\n"


def row_degrees(num_rows, average_degree, distribution, rng):
    if distribution == "uniform":
        lo, hi = average_degree // 2, int(average_degree * 1.5)
        return rng.integers(lo, hi + 1, num_rows).astype(np.int64)
    if distribution == "gamma":
        return np.maximum(rng.gamma(GAMMA_K, average_degree / GAMMA_K, num_rows).astype(np.int64), 1)
    raise ValueError(f"unknown distribution {distribution}")


def create_sparse_matrix(num_rows, max_cols, average_degree, distribution, seed=0, l2_norm=True):
    """Returns (x uint32 rows, y uint32 cols, val float64), row-sorted, columns sorted inside a row."""
    rng = np.random.default_rng(seed)
    deg = row_degrees(num_rows, average_degree, distribution, rng)
    total = int(deg.sum())
    x = np.repeat(np.arange(num_rows, dtype=np.int64), deg)
    y = rng.integers(0, max_cols, total, dtype=np.int64)
    key = np.sort(x * max_cols + y, kind="stable")   # sorted columns inside each row
    y = key % max_cols
    val = rng.random(total)
    if l2_norm:
        starts = np.concatenate([[0], np.cumsum(deg)[:-1]])
        norms = np.sqrt(np.add.reduceat(val * val, starts))
        val = val / np.repeat(norms, deg)
    return x.astype(np.uint32), y.astype(np.uint32), val


def matrix_name(num_rows, max_cols, average_degree, distribution):
    return f"matrix_{num_rows}_{max_cols}_{average_degree}_{distribution}.mtx"   # test_spmv_topk.py:104


def write_mtx(path, x, y, val, num_rows, max_cols, zero_indexed=False, precision=DEFAULT_PRECISION):
    base = 0 if zero_indexed else 1
    with open(path, "w") as f:
        f.write(MTX_HEADER.format(num_rows, max_cols, len(x)))
        chunk = 1 << 20
        for s in range(0, len(x), chunk):
            e = min(len(x), s + chunk)
            lines = [f"{int(a) + base} {int(b) + base} {v:.{precision}}\n" for a, b, v in zip(x[s:e], y[s:e], val[s:e])]
            f.writelines(lines)


def csr_from_coo(x, num_rows):
    counts = np.bincount(x, minlength=num_rows)
    ptr = np.zeros(num_rows + 1, np.uint64)
    np.cumsum(counts, out=ptr[1:])
    return ptr


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="create synthetic sparse matrices (MTX)")
    ap.add_argument("-r", "--rows", type=int, nargs="+", default=[10000])
    ap.add_argument("-c", "--cols", type=int, nargs="+", default=[512, 1024])
    ap.add_argument("-n", "--degree", type=int, nargs="+", default=[20, 40])
    ap.add_argument("-d", "--distribution", nargs="+", default=["uniform", "gamma"])
    ap.add_argument("-o", "--output", default="data/matrices_for_testing")
    ap.add_argument("-s", "--seed", type=int, default=0)
    ap.add_argument("-z", "--zero_index", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.output, exist_ok=True)
    for r in args.rows:
        for c in args.cols:
            for n in args.degree:
                for d in args.distribution:
                    x, y, v = create_sparse_matrix(r, c, n, d, seed=args.seed)
                    p = os.path.join(args.output, matrix_name(r, c, n, d))
                    write_mtx(p, x, y, v, r, c, zero_indexed=args.zero_index)
                    print(f"wrote {p}: {r}x{c}, nnz={len(x)}")
