"""CPU baseline -- the reference's test_cpu.py (:30-122) with the same CLI and CSV schema
(iter,rows,cols,nnz,K,exec_time_ms): MTX -> float64 scipy CSR -> per iteration a random L2-normalised query ->
Top-K SpMV on the host cores.

The reference calls sparse_dot_topn.awesome_cossim_topn(csr, vec.T, K, 0.0, use_threads=True, n_jobs=40)
(:104).  That package is not pinned by the reference and is absent from this image; when it cannot be
imported, the mathematically identical stand-in is used (and said so): float64 CSR x dense vector with
entries <= 0 dropped, then the K largest.  -j sets the thread count (the reference hard-codes 40)."""
import argparse
import os
import time
from datetime import datetime

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

try:   # pre-1.0 API used by the reference (test_cpu.py:14-20)
    from sparse_dot_topn import awesome_cossim_topn
    HAVE_SDT = True
except Exception:
    HAVE_SDT = False

INPUT_PATH = "data/matrices_for_testing/matrix_10000_1024_20_gamma.mtx"
NUM_TESTS = 30
OUTPUT_PATH = "data/results/cpu"
K = 100
THRESHOLD = 0.0


def load_mtx(path, zero_index):
    """test_cpu.py:63-88: the size line is line index 2 of the generator's 3-line header."""
    with open(path) as f:
        lines = f.readlines()
    rows, cols, size = (int(t) for t in lines[2].split(" ")[:3])
    data = np.loadtxt(lines[3:], dtype=np.float64, ndmin=2)
    base = 0 if zero_index else 1
    x = data[:, 0].astype(np.int64) - base
    y = data[:, 1].astype(np.int64) - base
    return rows, cols, size, csr_matrix((data[:, 2], (x, y)), shape=(rows, cols))


def topk_standin(csr, vec, k):
    y = csr @ vec
    y = np.where(y > THRESHOLD, y, 0.0)
    k = min(k, y.size)
    part = np.argpartition(-y, k - 1)[:k]
    order = part[np.lexsort((part, -y[part]))]
    return order, y[order]


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="run top-k spmv on cpu")
    ap.add_argument("-d", "--debug", action="store_true")
    ap.add_argument("-t", "--num_tests", type=int, default=NUM_TESTS)
    ap.add_argument("-k", "--K", type=int, default=K)
    ap.add_argument("-z", "--zero_index", action="store_true")
    ap.add_argument("-i", "--input", type=str, default=INPUT_PATH)
    ap.add_argument("-o", "--output", type=str, default=OUTPUT_PATH)
    ap.add_argument("-j", "--n_jobs", type=int, default=os.cpu_count())
    ap.add_argument("-s", "--seed", type=int, default=None)
    args = ap.parse_args()
    rows, cols, size, csr = load_mtx(args.input, args.zero_index)
    print(f"loaded matrix of size {rows}x{cols}, {size} nnz; backend="
          f"{'sparse_dot_topn' if HAVE_SDT else 'scipy float64 stand-in (sparse_dot_topn not installed)'}; cores={os.cpu_count()}")
    rng = np.random.default_rng(args.seed)
    results = []
    for t in range(args.num_tests):
        vec_np = rng.uniform(low=0.0, high=1.0, size=(cols,))
        vec_np /= np.linalg.norm(vec_np)
        start = time.time()
        if HAVE_SDT:
            res = awesome_cossim_topn(csr, csr_matrix(vec_np).transpose(), args.K, THRESHOLD, use_threads=True, n_jobs=args.n_jobs)
        else:
            res = topk_standin(csr, vec_np, args.K)
        end = (time.time() - start) * 1000
        results += [[t, rows, cols, size, args.K, end]]
        print(f"finished iteration {t + 1}/{args.num_tests}, time={end} ms" if args.debug else results[-1])
    df = pd.DataFrame(results, columns=["iter", "rows", "cols", "nnz", "K", "exec_time_ms"])
    out_path = args.output
    if not out_path.endswith(".csv"):
        out_dir = os.path.join(out_path, datetime.now().strftime("%Y_%m_%d_%H_%M_%S"))
        os.makedirs(out_dir, exist_ok=True)
        out_path = os.path.join(out_dir, os.path.splitext(os.path.basename(args.input))[0] + ".csv")
    else:
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    df.to_csv(out_path, index=False)
