"""Seeded input builders shared by the CPU oracle tests, the golden-vector script and the GPU parity tests.
Every builder returns (x uint32 rows, y uint32 cols, v float64 values, num_rows, num_cols), row-sorted."""
import numpy as np


def gamma(gen, rows=1500, cols=1024, deg=20, dist="gamma", seed=0):
    x, y, v = gen.create_sparse_matrix(rows, cols, deg, dist, seed=seed)
    return x, y, v, rows, cols


def massive_ties(rows=6000, seed=0):
    """Identical rows: every candidate ties, so the surviving indices depend on the exact slot dynamics of
    the replace-min lists (`>=` replacement, argmin -> highest slot among equal minima)."""
    per = 5
    x = np.repeat(np.arange(rows, dtype=np.uint32), per)
    y = np.tile(np.array([3, 99, 400, 401, 1000], np.uint32), rows)
    v = np.tile(np.array([0.5, 0.25, 0.125, 0.5, 0.3]), rows)
    rng = np.random.default_rng(seed)
    for r in rng.choice(rows, 40, replace=False):       # a few distinct rows sprinkled in
        v[r * per:(r + 1) * per] = rng.random(per)
    return x, y, v, rows, 1024


def long_rows(rows=640, seed=5, max_deg=400):
    """Rows of 1..max_deg non-zeros: most rows span many packets."""
    rng = np.random.default_rng(seed)
    deg = rng.integers(1, max_deg, rows)
    x = np.repeat(np.arange(rows, dtype=np.uint32), deg)
    y = rng.integers(0, 1024, x.size).astype(np.uint32)
    v = rng.random(x.size) / 20.0
    return x, y, v, rows, 1024


def one_row_per_partition(rows=32):
    x = np.arange(rows, dtype=np.uint32)
    return x, x.copy(), np.full(rows, 0.5), rows, 64


def make_query(cols, seed):
    """test_cpu.py:99-100 / utils.hpp:240-266: U[0,1)^C divided by its L2 norm, as float32."""
    rng = np.random.default_rng(seed)
    q = rng.random(cols)
    q /= np.linalg.norm(q)
    return q.astype(np.float32)
