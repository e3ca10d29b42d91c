import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tks():
    """The product package (loads libtopkspmv.so; builds it first if it is missing)."""
    from _pkg import pkg
    p = pkg()
    if not p.capi.LIB_PATH.exists():
        import importlib.util
        spec = importlib.util.spec_from_file_location("tks_build", ROOT / "approximate-spmv-topk_b200" / "build.py")
        b = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(b)
        b.build_lib()
    p.capi.lib()
    return p


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def gen(tks):
    return tks.create_matrices


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_required():
    if not has_cuda():
        pytest.fail("this test is marked gpu and needs a CUDA device (no CPU fallback exists)")


def make_query(cols, seed):
    """test_cpu.py:99-100 / utils.hpp:240-266: U[0,1)^C divided by its L2 norm, as float32."""
    rng = np.random.default_rng(seed)
    v = rng.random(cols)
    v /= np.linalg.norm(v)
    return v.astype(np.float32)


@pytest.fixture(scope="session")
def query():
    return make_query


def collect_from_workers(q, procs, n, timeout=300):
    """n items from the queue of spawned rank processes; fails at once (instead of after the whole timeout) when a
    worker has died -- its traceback is on stderr, which pytest shows with the failure."""
    import queue as _queue
    import time as _time
    got, t0 = [], _time.time()
    while len(got) < n:
        try:
            got.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError(f"a rank process died (exit codes {dead}); see captured stderr")
            if _time.time() - t0 > timeout:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError(f"only {len(got)} of {n} ranks reported within {timeout} s")
    return got
