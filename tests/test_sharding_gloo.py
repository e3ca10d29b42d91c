"""world_size-2 `gloo` test (CPU) of the multi-GPU plumbing: contiguous row shards, per-rank K
candidates as 64-bit ordering keys, all-gather, merge on every rank.  The per-rank scores come from a plain
float64 product here -- this test is about the exchange and the merge, not about the kernel."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, tie_higher, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _pkg import pkg
    tks = pkg()
    gen, sh = tks.create_matrices, tks.sharding
    rows, cols = 5000, 256
    x, y, v = gen.create_sparse_matrix(rows, cols, 12, "gamma", seed=3)
    ptr = gen.csr_from_coo(x, rows)
    rng = np.random.default_rng(9)
    vec = rng.random(cols)
    shards = sh.plan_row_shards_by_nnz(ptr, world)
    r0, r1 = shards[rank]
    p, idx, val = sh.slice_csr(ptr, y, v, r0, r1)
    # local scores of the shard (float32 like the engine's output), local row ids + row_offset
    deg = np.diff(p.astype(np.int64))
    lx = np.repeat(np.arange(r1 - r0), deg)
    score = np.zeros(r1 - r0)
    np.add.at(score, lx, val * vec[idx])
    score = score.astype(np.float32)
    order = np.lexsort((np.arange(score.size) if not tie_higher else -np.arange(score.size), -score.astype(np.float64)))[:k]
    mine = sh.make_keys(score[order], (order + r0).astype(np.uint32), tie_higher)
    mine = np.pad(mine, (0, k - mine.size))
    t = torch.from_numpy(mine.view(np.int64).copy())
    gathered = torch.empty(world * k, dtype=torch.int64)
    dist.all_gather_into_tensor(gathered, t)
    merged = sh.merge_topk_host([gathered.numpy().view(np.uint64)], k)
    ms, mr = sh.split_keys(merged, tie_higher)
    if rank == 0:
        q.put((ms, mr, shards))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("tie_higher", [False, True])
def test_two_rank_allgather_merge_equals_global_topk(tie_higher):
    world, k = 2, 50
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, tie_higher, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms, mr, shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # global reference
    sys.path.insert(0, str(ROOT))
    from _pkg import pkg
    gen = pkg().create_matrices
    x, y, v = gen.create_sparse_matrix(5000, 256, 12, "gamma", seed=3)
    vec = np.random.default_rng(9).random(256)
    score = np.zeros(5000)
    np.add.at(score, x, v * vec[y])
    score = score.astype(np.float32)
    order = np.lexsort((np.arange(5000) if not tie_higher else -np.arange(5000), -score.astype(np.float64)))[:k]
    assert np.array_equal(mr, order.astype(np.uint32))
    assert np.array_equal(ms, score[order])
    assert shards[0][1] == shards[1][0] and shards[0][0] == 0 and shards[1][1] == 5000


def _worker_batched(rank, world, port, k, batch, q):
    """ShardedSpMV.exchange under gloo: [batch, k] keys per rank -> [batch, world*k] regrouped by query."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _pkg import pkg
    tks = pkg()
    sh = tks.sharding
    rows, cols = 4000, 128
    rng = np.random.default_rng(1)
    scores = rng.random((batch, rows)).astype(np.float32)          # stand-in for the engine's row scores
    r0, r1 = sh.plan_row_shards_even(rows, world)[rank]
    mine = np.zeros((batch, k), np.uint64)
    for b in range(batch):
        loc = scores[b, r0:r1]
        order = np.lexsort((np.arange(loc.size), -loc.astype(np.float64)))[:k]
        mine[b] = sh.make_keys(loc[order], (order + r0).astype(np.uint32))
    s = tks.ShardedSpMV(engine=None, k=k, batch=batch)
    regrouped = s.exchange(torch.from_numpy(mine.view(np.int64).copy()))
    regrouped2 = s.exchange(torch.from_numpy(mine.view(np.int64).copy()))      # buffers are reused
    assert torch.equal(regrouped, regrouped2) and tuple(regrouped.shape) == (batch, world * k)
    merged = tks.distributed.merge_gathered_host(regrouped.numpy(), k)
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [1, 5])
def test_two_rank_batched_exchange_regroups_by_query(batch):
    world, k = 2, 20
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_batched, args=(r, world, port, k, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, str(ROOT))
    from _pkg import pkg
    sh = pkg().sharding
    scores = np.random.default_rng(1).random((batch, 4000)).astype(np.float32)
    for b in range(batch):
        order = np.lexsort((np.arange(4000), -scores[b].astype(np.float64)))[:k]
        ms, mr = sh.split_keys(merged[b])
        assert np.array_equal(mr, order.astype(np.uint32)) and np.array_equal(ms, scores[b][order])


# ---- FPGA mode: the P row partitions dealt out over the ranks, result words all-gathered, the reference's merge ----

class _OracleFixedEngine:
    """Stand-in for spmv.SpMVFixed under gloo (no GPU here): the oracle's literal kernel on the local partitions.
    Test infrastructure only -- it exercises ShardedSpMVFixed's shard planning, exchange and merge."""

    def __init__(self, x, y, val32, num_rows, num_cols, k=100, fixed_width=20, partitions=32, local_k=8,
                 limited_finished_rows=4, drift_free=False, device=0, device_pack=True):
        sys.path.insert(0, str(ROOT / "oracle"))
        import oracle
        self.o, self.Kp, self.LFR, self.df = oracle, local_k, limited_finished_rows, drift_free
        self.packed = oracle.pack_bscsr(x, y, val32, num_rows, partitions, fixed_width)

    def first_row_array(self):
        return self.packed["first_row"]

    def reset(self, vec32):
        self.vec32 = np.asarray(vec32, np.uint32)

    def __call__(self):
        self.words = self.o.bscsr_kernel(self.packed, self.vec32, self.Kp, self.LFR, self.df)
        return 0

    def read_partition_results(self):
        return self.words

    def close(self):
        pass


def _worker_fixed(rank, world, port, rows, P, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from _pkg import pkg
    tks = pkg()
    x, y, v = tks.create_matrices.create_sparse_matrix(rows, 1024, 12, "gamma", seed=4)
    val32 = oracle.fx32_from_double(v)
    rng = np.random.default_rng(2)
    vec = rng.random(1024); vec = (vec / np.linalg.norm(vec)).astype(np.float32)
    s = tks.ShardedSpMVFixed(x, y, val32, rows, 1024, k=100, partitions=P, engine_factory=_OracleFixedEngine)
    s.reset(oracle.query_fx32_from_f32(vec))
    s()
    val, idx = s.read_result()
    # submit / fetch (without NCCL: the blocking verbs behind the same tickets): same lists, two queries kept
    vec2 = rng.random(1024); vec2 = (vec2 / np.linalg.norm(vec2)).astype(np.float32)
    t1 = s.submit(oracle.query_fx32_from_f32(vec))
    t2 = s.submit(oracle.query_fx32_from_f32(vec2))
    v1, i1 = s.fetch(t1)
    assert np.array_equal(v1, val) and np.array_equal(i1, idx)
    v2, i2 = s.fetch(t2)
    s.reset(oracle.query_fx32_from_f32(vec2)); s()
    vb, ib = s.read_result()
    assert np.array_equal(v2, vb) and np.array_equal(i2, ib)
    try:
        s.fetch(t1)
        raise AssertionError("a ticket can be fetched once")
    except tks.capi.TksError:
        pass
    q.put((rank, val.copy(), idx.copy(), s.first_row.copy(), (s.r0, s.r1)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("rows,P", [(8000, 32), (7777, 8)])
def test_two_rank_fixed_mode_partitions_equal_one_device(rows, P):
    """Partitions dealt out over 2 ranks (incl. a row count that does not divide: the last shard is padded with virtual
    rows so that rows_per_part stays the unsharded one) == all partitions on one device, bit for bit."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fixed, args=(r, world, port, rows, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, val, idx, first_row, shard = q.get(timeout=180)
        got[r] = (val, idx, first_row, shard)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    from _pkg import pkg
    x, y, v = pkg().create_matrices.create_sparse_matrix(rows, 1024, 12, "gamma", seed=4)
    rng = np.random.default_rng(2)
    vec = rng.random(1024); vec = (vec / np.linalg.norm(vec)).astype(np.float32)
    o = oracle.bscsr_topk(x, y, v, rows, vec, P=P)
    for r in range(world):
        val, idx, first_row, shard = got[r]
        assert np.array_equal(first_row, o["packed"]["first_row"])
        assert np.array_equal(idx, o["idx"][:100]) and np.array_equal(val, o["val"][:100])
    assert got[0][3][1] == got[1][3][0]
