"""Accuracy metrics of an approximate top-k list against the exact one ("top-K recall" half of the metric).

Definitions follow the reference's own analysis scripts so that numbers are comparable with its figures:
  precision   |approx[:t] & exact[:t]| / t          host_spmv_bscsr.cpp:646-650, plot_errors.py:85-93, 190-193
  kendall_tau pairwise rank agreement over the union  plot_errors.py:197-233
  ndcg        relevance = exact score                 plot_errors.py:236-247
  closed form / Monte-Carlo model of the precision lost by keeping only `partition_k` candidates in each of
  `b` row partitions                                  src/resources/python/topk_errors.py:29-42, 47-83
Plain Python/NumPy on k-element lists: this is reporting code, not part of the hot path.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

THRESHOLDS = (8, 16, 32, 50, 75, 100)      # plot_errors.py:36


def precision_at(exact_idx, approx_idx, t=None):
    """|exact[:t] & approx[:t]| / t  (the reference's "precision"; with t = k it is the top-K recall)."""
    t = len(exact_idx) if t is None else t
    return len(set(map(int, exact_idx[:t])) & set(map(int, approx_idx[:t]))) / t


def kendall_tau(reference_rank, predicted_rank):
    """(concordant - discordant pairs) / sqrt(pairs ranked by the reference * pairs ranked by the prediction)."""
    ref = {int(item): pos for pos, item in enumerate(reference_rank)}
    pred = {int(item): pos for pos, item in enumerate(predicted_rank)}
    items = list(set(ref) | set(pred))
    c_plus = c_minus = c_s = c_u = 0
    for i in range(len(items)):
        for j in range(i + 1, len(items)):
            a, b = items[i], items[j]
            in_ref = a in ref and b in ref
            in_pred = a in pred and b in pred
            c_u += in_ref
            c_s += in_pred
            if in_ref and in_pred:
                if (ref[a] - ref[b]) * (pred[a] - pred[b]) > 0:
                    c_plus += 1
                else:
                    c_minus += 1
    return (c_plus - c_minus) / (math.sqrt(c_u) * math.sqrt(c_s))


def ndcg(exact_idx, exact_val, approx_idx, approx_val=None):
    """DCG of the approximate list with the exact scores as relevance, over the ideal DCG."""
    rel = {int(i): float(v) for i, v in zip(exact_idx, exact_val)}
    dcg = sum(rel.get(int(idx), 0.0) / math.log2(i + 2) for i, idx in enumerate(approx_idx))
    idcg = sum(float(v) / math.log2(i + 2) for i, v in enumerate(exact_val))
    return dcg / idcg


def closed_form_approx(n, b, k, partition_k):
    """Probability model of topk_errors.py:29-39: n rows in b partitions, partition_k kept per partition."""
    if k <= partition_k:
        return 1
    if partition_k * b < k:
        return 0
    denom = math.comb(n, k)
    delta = 0
    for i in range(partition_k + 1, min(n // b, k)):
        delta += math.comb(n // b, i)
    return 1 - Fraction(b * delta, denom)


def closed_form_precision_estimation(n, b, k, partition_k):
    return float(np.mean([closed_form_approx(n, b, k_i, partition_k) for k_i in range(1, k + 1)]))


def monte_carlo_partition_precision(n, b, k, partition_k, trials=10, seed=0):
    """Expected precision when the top-k is taken from the union of the per-partition top-`partition_k` of
    i.i.d. uniform scores (topk_errors.py:47-83, seeded)."""
    rng = np.random.default_rng(seed)
    starts = [i * (n // b) + min(i, n % b) for i in range(b)]
    out = []
    for _ in range(trials):
        scores = rng.random(n)
        exact = np.argsort(scores)[-k:]
        cand = np.concatenate([np.argsort(part)[-partition_k:] + starts[i]
                               for i, part in enumerate(np.array_split(scores, b))])
        approx = cand[np.argsort(scores[cand])[-k:]]
        out.append(len(set(exact.tolist()) & set(approx.tolist())) / k)
    return float(np.mean(out))


def report(exact_idx, exact_val, approx_idx, approx_val=None, thresholds=THRESHOLDS):
    """All metrics at the reference's thresholds, as a flat dict (bench.py / the sweep CSV append these)."""
    out = {}
    for t in thresholds:
        if t > len(exact_idx) or t > len(approx_idx):
            continue
        out[f"precision@{t}"] = precision_at(exact_idx, approx_idx, t)
        out[f"kendall_tau@{t}"] = kendall_tau(list(exact_idx[:t]), list(approx_idx[:t]))
        out[f"ndcg@{t}"] = ndcg(exact_idx[:t], exact_val[:t], approx_idx[:t])
    return out
