"""CPU tests of the host-side checkers bench.py relies on for `parity_n` and the cfg3 accuracy record: they decide whether
a multi-GPU line is accepted, so they are held to known answers here."""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_merge_lists_is_the_global_topk_under_the_stated_order():
    b = _bench()
    rng = np.random.default_rng(0)
    scores = rng.random(4000).astype(np.float32)
    scores[[10, 3000]] = scores.max() + 1.0                      # an exact tie across two ranks
    lists = []
    for r0, r1 in ((0, 1000), (1000, 2500), (2500, 4000)):
        loc = scores[r0:r1]
        order = np.lexsort((np.arange(loc.size), -loc.astype(np.float64)))[:50]
        lists.append((order + r0, loc[order]))
    gi, gv = b.merge_lists(lists, 50)
    want = np.lexsort((np.arange(4000), -scores.astype(np.float64)))[:50]
    assert np.array_equal(gi, want) and np.array_equal(gv, scores[want])
    assert gi[0] == 10 and gi[1] == 3000                        # tie: lower row first


def test_lists_agree_accepts_near_ties_and_rejects_real_differences():
    b = _bench()
    gv = np.linspace(1.0, 0.5, 100).astype(np.float32)
    gi = np.arange(100)
    ok, _ = b.lists_agree(gi, gv, gi, gv)
    assert ok
    # the k-th place swapped with a row whose score differs by 1e-7 relative: accepted
    idx = gi.copy(); idx[-1] = 777
    val = gv.copy(); val[-1] = gv[-1] * (1 + 5e-8)
    ok, d = b.lists_agree(idx, val, gi, gv)
    assert ok and d["set_difference"] == 2
    # a row from the middle of the list missing: rejected
    idx = gi.copy(); idx[40] = 778
    ok, _ = b.lists_agree(idx, gv, gi, gv)
    assert not ok
    # a score off by 1e-3 relative: rejected
    val = gv.copy(); val[5] *= 1.001
    ok, _ = b.lists_agree(gi, val, gi, gv)
    assert not ok
    ok, _ = b.lists_agree(gi[:99], gv[:99], gi, gv)
    assert not ok


def test_lfr_overflow_packet_count_on_a_hand_made_stream():
    b = _bench()
    # one partition, B = 15: packet 0 holds rows of 3 non-zeros (5 segments > LFR 4), packet 1 one long row, packet 2 = tail
    x = np.concatenate([np.repeat(np.arange(5), 3), np.full(15, 5), np.repeat([6, 7], [2, 3])]).astype(np.uint32)
    r = b.lfr_overflow_packets(x, 8, 1, 15, 4)
    assert r["packets"] == 1 and r["of"] == 3 and r["partitions_hit"] == 1 and r["mean_position_of_first_event_in_its_partition"] == 0.0
    r2 = b.lfr_overflow_packets(x, 8, 1, 15, 5)
    assert r2["packets"] == 0 and r2["partitions_hit"] == 0


def test_sharded_submit_host_refuses_what_it_cannot_pipeline():
    sys.path.insert(0, str(ROOT))
    from _pkg import pkg
    import pytest
    tks = pkg()
    s = tks.ShardedSpMV(engine=None, k=10, batch=4)           # no process group: one rank
    with pytest.raises(ValueError):
        s.submit_host(np.zeros(8, np.float32))
    with pytest.raises(ValueError):
        s.submit(0)
