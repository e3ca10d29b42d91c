#!/usr/bin/env python
"""Generates tests/golden/*.npz: outputs of the REFERENCE ITSELF on small seeded inputs.

Run in the build container, where /root/reference exists and `make -C oracle ref` has produced
  oracle/_ref/libref_gold.so            the reference's gold_algorithms.hpp / evaluation_utils.hpp (float path)
  oracle/_ref/libref_fpga_w*_k*_l*_p*.so the reference's FPGA host (host_spmv_bscsr.cpp) and HLS kernel
                                        (spmv_bscsr_top_k_multicore.{hpp,cpp}) compiled against oracle/shim
The reference ships no golden vectors or known-answer tests, and /root/reference does not exist on
the GPU box, so these fixtures are what pins the oracle (and through it the CUDA path) there.  Each file holds
the INPUTS (so nothing depends on a NumPy generator staying stable) and the reference's outputs.

    python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
import oracle  # noqa: E402
from _pkg import pkg  # noqa: E402

gen = pkg().create_matrices

# name -> (builder, query seed)
MATRICES = {
    "gamma20": (lambda: cases.gamma(gen, 1536, 1024, 20, "gamma", seed=0), 1),
    "uniform40": (lambda: cases.gamma(gen, 800, 1024, 40, "uniform", seed=3), 2),
    "short_rows": (lambda: cases.gamma(gen, 4000, 1024, 2, "gamma", seed=2), 3),      # > LFR segments per packet: row-counter drift
    "ties": (lambda: cases.massive_ties(1200, seed=0), 8),
    "long_rows": (lambda: cases.long_rows(160, seed=5), 9),
}
# (matrix, W, Kp, LFR, P): every combination has a library in oracle/Makefile REF_FPGA_COMBOS
FIXED = [("gamma20", 20, 8, 4, 32), ("gamma20", 21, 8, 4, 32), ("gamma20", 25, 8, 4, 32), ("gamma20", 26, 8, 4, 32),
         ("gamma20", 32, 8, 4, 32), ("gamma20", 20, 1, 4, 32), ("gamma20", 20, 2, 4, 32), ("gamma20", 20, 4, 4, 32),
         ("gamma20", 20, 8, 1, 32), ("gamma20", 20, 8, 2, 32), ("gamma20", 20, 8, 3, 32), ("gamma20", 20, 8, 4, 4),
         ("gamma20", 20, 8, 4, 64), ("gamma20", 32, 4, 2, 8),
         ("uniform40", 20, 8, 4, 32), ("uniform40", 32, 8, 4, 32),
         ("short_rows", 20, 8, 4, 32), ("short_rows", 32, 8, 4, 32), ("short_rows", 20, 8, 2, 32), ("short_rows", 20, 8, 4, 4),
         ("ties", 20, 8, 4, 32), ("ties", 20, 4, 4, 32), ("ties", 32, 4, 2, 8), ("ties", 20, 8, 4, 4),
         ("long_rows", 20, 8, 4, 4), ("long_rows", 32, 4, 2, 8)]
FLOAT_K = [1, 8, 100]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert oracle.ref_gold() is not None, "oracle/_ref/libref_gold.so missing: run `make -C oracle ref`"
    for name, (build, qseed) in MATRICES.items():
        x, y, v, rows, cols = build()
        vec = cases.make_query(cols, qseed)
        val32 = oracle.fx32_from_double(v)
        vec32 = oracle.query_fx32_from_f32(vec)
        out = dict(x=x, y=y, v=v, rows=rows, cols=cols, vec=vec, val32=val32, vec32=vec32)
        # float path: the reference's spmv_coo_gold_top_k (+ sort_tuples), gold_algorithms.hpp:188-246
        v32 = v.astype(np.float32)
        for k in FLOAT_K:
            ui, uv = oracle.ref_gold_topk_f32(x, y, v32, vec, k, sort=False)
            si, sv = oracle.ref_gold_topk_f32(x, y, v32, vec, k, sort=True)
            out.update({f"gold_k{k}_slots_idx": ui, f"gold_k{k}_slots_val": uv, f"gold_k{k}_idx": si, f"gold_k{k}_val": sv})
        np.savez_compressed(HERE / f"inputs_{name}.npz", **out)
        print(f"inputs_{name}.npz: {rows}x{cols}, nnz={x.size}")
    for name, W, Kp, LFR, P in FIXED:
        d = np.load(HERE / f"inputs_{name}.npz")
        R = oracle.ref_fpga(W, Kp, LFR, P)
        assert R is not None, f"missing {oracle.ref_fpga_path(W, Kp, LFR, P)}"
        # the reference's own quantisation of a few values, to pin (T) double and write_block_val
        probe = np.linspace(0.0, 1.0, 257)[:-1] ** 2
        q32 = np.array([R.ref_fx32_from_double(float(t)) for t in probe], np.uint32)
        qW = np.array([R.ref_fxW_from_fx32(int(t)) for t in q32], np.uint32)
        ref = oracle.RefFpga(d["x"], d["y"], d["val32"], int(d["rows"]), int(d["cols"]), d["vec32"], W, Kp, LFR, P)
        info = ref.partition_info()
        packets = np.concatenate(ref.packets(), axis=0)
        qblocks = ref.query_blocks()
        ref.run()
        iw, vw = ref.result_words()
        ri, rv = ref.read_result()
        gi, gv = np.zeros(100, np.uint32), np.zeros(100, np.uint32)
        R.ref_gold_topk_fx32(np.ascontiguousarray(d["x"]), np.ascontiguousarray(d["y"]), np.ascontiguousarray(d["val32"]),
                             d["x"].size, np.ascontiguousarray(d["vec32"]), int(d["cols"]), 100, gi, gv)
        ref.close()
        B = R.packet_size
        np.savez_compressed(HERE / f"fixed_{name}_w{W}_k{Kp}_l{LFR}_p{P}.npz", W=W, Kp=Kp, LFR=LFR, P=P, B=B,
                            probe=probe, probe_fx32=q32, probe_fxW=qW,
                            first_row=info[:, 0], last_row=info[:, 1], nnz_per_part=info[:, 2], packets_per_part=info[:, 3],
                            packets_sha256=sha(packets), packets_head=packets[:64], query_blocks=qblocks,
                            # positions >= B of the result words are uninitialised struct padding in the reference
                            idx_words=iw[:, :, :B], val_words=vw[:, :, :B], merged_idx=ri, merged_val=rv,
                            gold_fx32_idx=gi, gold_fx32_val=gv)
        print(f"fixed_{name}_w{W}_k{Kp}_l{LFR}_p{P}.npz: {packets.shape[0]} packets, {ri.size} merged results")


if __name__ == "__main__":
    main()
