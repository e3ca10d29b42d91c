"""Sweep driver -- same grid, command templates and output naming as the reference's test_spmv_topk.py
(:12-111), with a new test kind "b200" that runs this repo's host executable (the drop-in for the FPGA /
GPU hosts).  Output files: {kind}_{rows}_{cols}_{dist}_{nnz}_{design}_{K}_{NITER}.csv in OUT_FOLDER.

    python test_spmv_topk.py [--tests b200 cpu] [--matrix-folder DIR] [--sizes 10000 100000] ...

Matrices follow the reference naming matrix_{rows}_{cols}_{nnz}_{dist}.mtx (:104); missing ones are
generated with approximate-spmv-topk_b200/create_matrices.py (1-indexed, like the reference generator).
"""
import argparse
import os
import subprocess
import sys
from datetime import datetime
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

##########################################################
# Configuration (reference defaults, test_spmv_topk.py:12-21)
##########################################################
TESTS = ["b200"]
DEBUG = False
MATRIX_SIZES = [5_000_000, 10_000_000, 15_000_000]
MATRIX_COLS = [512, 1024]
MATRIX_DIST = ["uniform", "gamma"]
MATRIX_NNZ = [20, 40]
K = 100
NITER = 30
ZERO_INDEXED = False   # the generator writes 1-indexed files (create_matrices.py:120,124)

# B200 designs: (name, extra flags).  They mirror the FPGA builds of test_spmv_topk.py:41-47
# (32/26/21-bit fixed, float) plus the paper's 20-bit design and the GPU sweep's half-precision variant
# (GPU_USE_HALF, test_spmv_topk.py:54), all from ONE binary.
B200_DESIGNS = [("float", ""), ("half", "-a"), ("20bit", "-f -w 20"), ("21bit", "-f -w 21"), ("26bit", "-f -w 26"), ("32bit", "-f -w 32")]
B200_EXE = str(ROOT / "build" / "topk-spmv-b200")

B200_CMD = "{} {} -t {} -m {} -k {} {} {} | tee {}"
CPU_CMD = "{} test_cpu.py {} -t {} {} -i {} -k {} -o {}"


def test(t, s, c, d, n, input_matrix, out_folder, designs, niter, k):
    results = []
    if t == "b200":
        for name, flags in designs:
            output_file = os.path.join(out_folder, f"{t}_{s}_{c}_{d}_{n}_{name}_{k}_{niter}.csv")
            cmd = B200_CMD.format(B200_EXE, "-d" if DEBUG else "", niter, input_matrix, k, "-z" if ZERO_INDEXED else "", flags, output_file)
            print(f"running {cmd}", flush=True)
            # `| tee` (kept from the reference's command templates, test_spmv_topk.py:62-64) would make the shell report
            # tee's status: pipefail keeps the executable's, so that a failing run is counted
            results.append(subprocess.run(["bash", "-o", "pipefail", "-c", cmd], stdout=subprocess.PIPE, stderr=subprocess.PIPE))
    elif t == "cpu":
        output_file = os.path.join(out_folder, f"{t}_{s}_{c}_{d}_{n}_{k}_{niter}.csv")
        cmd = CPU_CMD.format(sys.executable, "-d" if DEBUG else "", niter, "-z" if ZERO_INDEXED else "", input_matrix, k, output_file)
        print(f"running {cmd}", flush=True)
        results.append(subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE))
    else:
        raise ValueError(f"unknown test kind {t} (the fpga/gpu kinds belong to the reference's own hosts)")
    return results


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tests", nargs="+", default=TESTS)
    ap.add_argument("--matrix-folder", default=str(ROOT / "data" / "matrices_for_testing"))
    ap.add_argument("--out-folder", default=str(ROOT / "data" / "results" / datetime.now().strftime("%Y_%m_%d_%H_%M_%S")))
    ap.add_argument("--sizes", type=int, nargs="+", default=MATRIX_SIZES)
    ap.add_argument("--cols", type=int, nargs="+", default=MATRIX_COLS)
    ap.add_argument("--dist", nargs="+", default=MATRIX_DIST)
    ap.add_argument("--nnz", type=int, nargs="+", default=MATRIX_NNZ)
    ap.add_argument("--designs", nargs="+", default=[d[0] for d in B200_DESIGNS])
    ap.add_argument("-k", type=int, default=K)
    ap.add_argument("-t", "--niter", type=int, default=NITER)
    args = ap.parse_args()
    os.makedirs(args.out_folder, exist_ok=True)
    os.makedirs(args.matrix_folder, exist_ok=True)
    designs = [d for d in B200_DESIGNS if d[0] in args.designs]
    from _pkg import pkg
    gen = pkg().create_matrices
    failures = 0
    for t in args.tests:
        for s in args.sizes:
            for c in args.cols:
                for d in args.dist:
                    for n in args.nnz:
                        input_matrix = os.path.join(args.matrix_folder, gen.matrix_name(s, c, n, d))
                        if not os.path.exists(input_matrix):
                            x, y, v = gen.create_sparse_matrix(s, c, n, d, seed=0)
                            gen.write_mtx(input_matrix, x, y, v, s, c, zero_indexed=ZERO_INDEXED)
                        for r in test(t, s, c, d, n, input_matrix, args.out_folder, designs, args.niter, args.k):
                            if r.returncode != 0:
                                failures += 1
                                print(r.stderr.decode(errors="replace")[-2000:], file=sys.stderr)
    sys.exit(1 if failures else 0)
