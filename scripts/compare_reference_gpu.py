"""Same-box comparison with the reference's own GPU host.

The reference's `src/gpu/host_spmv_topk_csr_gpu.cu` (cuSPARSE or LightSpMV SpMV -> N-vector -> thrust sort ->
get_topk) is compiled UNMODIFIED by `make -C oracle ref_gpu` into `oracle/_ref/ref_gpu_csr_topk` (one -D for an
enum cuSPARSE 12 renamed).  This script writes one synthetic MTX file of the benchmark law, runs that binary
(-i 0 cuSPARSE, -i 1 LightSpMV, each with and without -a half precision) and this repo's drop-in host
executable `build/topk-spmv-b200` on the same file, and prints one JSON object with the mean per-query times
each program reports in its own CSV (`hw_exec_time_ms`: kernel + synchronisation, the first two iterations
skipped like the reference's `mean(x, 2)`, host_spmv_bscsr.cpp:699).

    python scripts/compare_reference_gpu.py --rows 2000000 --iters 12 > gpurun_out/compare_ref_gpu.json

It is test/measurement infrastructure: nothing in the product path imports it.  The reference binary only exists
where /root/reference was present at build time (this container); on the GPU box the prebuilt file is used.
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF_EXE = ROOT / "oracle" / "_ref" / "ref_gpu_csr_topk"
OUR_EXE = ROOT / "build" / "topk-spmv-b200"


def write_mtx_fast(path, x, y, val, rows, cols):
    """0-indexed MTX (what the reference GPU host expects, host_spmv_topk_csr_gpu.cu:324) through pyarrow's CSV writer."""
    import pyarrow as pa
    import pyarrow.csv as pacsv
    with open(path, "wb") as f:
        f.write(f"%%MatrixMarket matrix coordinate real general\n%\n{rows} {cols} {len(x)}\n".encode())
        tbl = pa.table({"r": pa.array(x), "c": pa.array(y), "v": pa.array(np.round(val, 10))})
        pacsv.write_csv(tbl, f, pacsv.WriteOptions(include_header=False, delimiter=" ", quoting_style="none"))


def run_csv(cmd, timeout):
    t0 = time.time()
    out = subprocess.run(list(map(str, cmd)), capture_output=True, text=True, timeout=timeout)
    wall = time.time() - t0
    if out.returncode != 0:
        return {"error": out.stderr[-500:], "returncode": out.returncode}
    text = out.stdout
    start = text.find("iteration,")
    rows = list(csv.DictReader(io.StringIO(text[start:]))) if start >= 0 else []
    if not rows:
        return {"error": "no CSV rows", "stdout_tail": text[-300:]}
    skip = 2 if len(rows) > 4 else 0

    def mean(col):
        v = [float(r[col]) for r in rows[skip:] if r.get(col) not in (None, "")]
        return float(np.mean(v)) if v else None

    res = {"iterations": len(rows), "wall_s": round(wall, 2),
           "hw_exec_time_ms": mean("hw_exec_time_ms"), "hw_spmv_only_time_ms": mean("hw_spmv_only_time_ms"),
           "readback_time_ms": mean("readback_time_ms"), "hw_setup_time_ms": mean("hw_setup_time_ms"),
           "error_idx_mean": mean("error_idx"), "error_val_mean": mean("error_val")}
    if "precision" in rows[0]:
        res["precision"] = mean("precision")
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--cols", type=int, default=1024)
    ap.add_argument("--degree", type=int, default=20)
    ap.add_argument("--dist", default="gamma")
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("-k", type=int, default=100)
    ap.add_argument("--dir", default="/tmp/tks_compare")
    ap.add_argument("--timeout", type=int, default=900)
    args = ap.parse_args()
    from _pkg import pkg
    gen = pkg().create_matrices
    os.makedirs(args.dir, exist_ok=True)
    path = Path(args.dir) / gen.matrix_name(args.rows, args.cols, args.degree, args.dist)
    t0 = time.time()
    x, y, v = gen.create_sparse_matrix(args.rows, args.cols, args.degree, args.dist, seed=0)
    write_mtx_fast(path, x, y, v, args.rows, args.cols)
    res = {"matrix": path.name, "rows": args.rows, "cols": args.cols, "nnz": int(len(x)), "k": args.k,
           "mtx_bytes": path.stat().st_size, "generate_write_s": round(time.time() - t0, 1), "runs": {}}
    nnz = len(x)
    del x, y, v
    if REF_EXE.exists():
        for name, flags in [("reference_cusparse_fp32", ["-i", 0]), ("reference_lightspmv_fp32", ["-i", 1]),
                            ("reference_cusparse_fp16", ["-i", 0, "-a"]), ("reference_lightspmv_fp16", ["-i", 1, "-a"])]:
            res["runs"][name] = run_csv([REF_EXE, "-m", path, "-k", args.k, "-t", args.iters, *flags], args.timeout)
    else:
        res["runs"]["reference"] = {"error": f"{REF_EXE} not built (needs /root/reference at build time)"}
    for name, flags in [("b200_fp32", []), ("b200_fp16", ["-a"]), ("b200_fixed20_driftfree", ["-f", "-w", 20, "-D"])]:
        res["runs"][name] = run_csv([OUR_EXE, "-m", path, "-z", "-k", args.k, "-t", args.iters, "-e", 1, *flags],
                                    args.timeout)
    for r in res["runs"].values():
        if r.get("hw_exec_time_ms"):
            r["nnz_per_s"] = nnz / (r["hw_exec_time_ms"] * 1e-3)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
