"""Golden values of the accuracy metrics from the reference's OWN functions.

plot_errors.py needs matplotlib/seaborn and topk_errors.py runs a long loop at import, so the function
definitions are lifted out of the files with `ast` (unmodified source) and executed on seeded lists.
Run here (where /root/reference exists):   python tests/golden/make_accuracy_golden.py
"""
import ast
import json
import math
from fractions import Fraction
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src/resources/python")


def lift(path, names):
    tree = ast.parse(path.read_text())
    ns = {"np": np, "math": math, "Fraction": Fraction}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), str(path), "exec"), ns)
    return ns


def main():
    pe = lift(REF / "plotting" / "plot_errors.py", {"kendall_tau", "ndcg"})
    te = lift(REF / "topk_errors.py", {"closed_form_approx", "closed_form_precision_estimation"})
    rng = np.random.default_rng(0)
    cases = []
    for t in (8, 16, 50, 100):
        exact_idx = rng.choice(100000, t, replace=False).tolist()
        exact_val = np.sort(rng.random(t))[::-1].tolist()
        approx_idx = list(exact_idx)
        for _ in range(max(1, t // 8)):                       # a few swaps and a few foreign rows
            i, j = rng.integers(0, t, 2)
            approx_idx[i], approx_idx[j] = approx_idx[j], approx_idx[i]
        for i in rng.choice(t, max(1, t // 16), replace=False):
            approx_idx[i] = int(200000 + i)
        approx_val = [0.0] * t
        cases.append({"exact_idx": exact_idx, "exact_val": exact_val, "approx_idx": approx_idx,
                      "kendall_tau": float(pe["kendall_tau"](exact_idx, approx_idx)),
                      "ndcg": float(pe["ndcg"](exact_idx, exact_val, approx_idx, approx_val)[0]),
                      "precision": len(set(exact_idx) & set(approx_idx)) / t})
    closed = []
    for n, b, k, pk in [(100000, 8, 8, 8), (100000, 8, 16, 8), (100000, 16, 100, 8), (100000, 32, 100, 8),
                        (100000, 32, 300, 8), (100000, 8, 100, 8), (1000, 4, 50, 8)]:
        closed.append({"n": n, "b": b, "k": k, "partition_k": pk,
                       "approx": float(te["closed_form_approx"](n, b, k, pk)),
                       "precision": float(te["closed_form_precision_estimation"](n, b, k, pk))})
    out = Path(__file__).with_name("accuracy_golden.json")
    out.write_text(json.dumps({"cases": cases, "closed_form": closed}, indent=1))
    print("wrote", out)


if __name__ == "__main__":
    main()
