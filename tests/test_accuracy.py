"""The accuracy metrics (precision / Kendall tau / NDCG, partition-local-K model) against golden values
produced by the reference's own functions (tests/golden/make_accuracy_golden.py)."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).parent / "golden" / "accuracy_golden.json").read_text())


@pytest.fixture(scope="module")
def acc(tks):
    return tks.accuracy


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"t{len(c['exact_idx'])}")
def test_rank_metrics_match_reference_functions(acc, case):
    e, ev, a = case["exact_idx"], case["exact_val"], case["approx_idx"]
    assert acc.precision_at(e, a) == case["precision"]
    assert acc.kendall_tau(e, a) == pytest.approx(case["kendall_tau"], rel=1e-12)
    assert acc.ndcg(e, ev, a) == pytest.approx(case["ndcg"], rel=1e-12)
    assert acc.kendall_tau(e, e) == pytest.approx(1.0) and acc.ndcg(e, ev, e) == pytest.approx(1.0)


@pytest.mark.parametrize("c", GOLD["closed_form"], ids=lambda c: f"n{c['n']}_b{c['b']}_k{c['k']}")
def test_closed_form_model_matches_reference(acc, c):
    assert float(acc.closed_form_approx(c["n"], c["b"], c["k"], c["partition_k"])) == pytest.approx(c["approx"], rel=1e-12)
    assert acc.closed_form_precision_estimation(c["n"], c["b"], c["k"], c["partition_k"]) == pytest.approx(c["precision"], rel=1e-12)


def test_monte_carlo_agrees_with_the_model_where_it_is_exact(acc):
    # k <= partition_k: nothing can be lost
    assert acc.monte_carlo_partition_precision(20000, 8, 8, 8, trials=3) == 1.0
    # 32 partitions x 8 candidates, k = 100: a small, non-zero loss (the paper's operating point)
    p = acc.monte_carlo_partition_precision(100000, 32, 100, 8, trials=20, seed=1)
    assert 0.9 < p <= 1.0
    # 8 partitions x 8 candidates cannot hold a top-100: at most 64 of them survive
    assert acc.monte_carlo_partition_precision(100000, 8, 100, 8, trials=3) <= 0.64


def test_report_keys(acc):
    rng = np.random.default_rng(0)
    e = rng.choice(10000, 100, replace=False)
    ev = np.sort(rng.random(100))[::-1]
    r = acc.report(e, ev, e)
    assert r["precision@100"] == 1.0 and r["kendall_tau@8"] == pytest.approx(1.0) and set(r) >= {"ndcg@50", "precision@16"}
