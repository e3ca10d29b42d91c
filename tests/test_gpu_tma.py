"""The bulk-copy variant of the main kernel (TKS_TMA=1: per-warp rings of cp.async.bulk + mbarrier instead of register
loads; measured slower and therefore off by default) must give the results of the default path.  The switch
is read once per process, so the check runs in a child process."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

CHILD = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, oracle
from _pkg import pkg
tks = pkg(); gen = tks.create_matrices
rows = 150000
x, y, v = gen.create_sparse_matrix(rows, 1024, 20, "gamma", seed=3)
keep = (x %% 11) != 5                      # some empty rows
x, y, v = x[keep], y[keep], v[keep].astype(np.float32)
ptr = gen.csr_from_coo(x, rows)
for kw in ({}, {"half": True}, {"bf16": True}):
    with tks.SpMV(ptr, y, v, rows, 1024, k=100, **kw) as s:
        for seed in (1, 2, 3):
            r = np.random.default_rng(seed).random(1024); q = (r / np.linalg.norm(r)).astype(np.float32)
            s.reset(q); s(); val, idx, cnt = s.read_result()
            gold = oracle.gold_topk_f16 if kw.get("half") else oracle.gold_topk_bf16 if kw.get("bf16") else oracle.gold_topk_f32
            gi, gv = gold(x, y, v, q, 100)
            assert cnt == 100
            np.testing.assert_allclose(val, gv, rtol=1e-5)
            assert len(set(idx.tolist()) ^ set(gi.tolist())) <= 2
            t = s.submit_host(q, 100); pv, pi, pc = s.fetch(t)          # and through the pipeline
            assert np.array_equal(pi, idx) and np.array_equal(pv, val)
print("tma ok")
"""


def test_bulk_copy_variant_matches_the_oracle(cuda_required):
    import os
    env = dict(os.environ, TKS_TMA="1")
    out = subprocess.run([sys.executable, "-c", CHILD % (str(ROOT), str(ROOT / "oracle"))], capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0 and "tma ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
