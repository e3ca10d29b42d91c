"""One process per GPU over NCCL: contiguous row shards, all-gather of the K candidates, merge on
every rank.  Uses 2 ranks when the box has >= 2 GPUs, else a single rank through the same code path."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, q, exchange="auto", k=100):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from _pkg import pkg
    from conftest import make_query
    tks = pkg()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    gen, sh = tks.create_matrices, tks.sharding
    rows, cols = 20000, 1024
    x, y, v = gen.create_sparse_matrix(rows, cols, 20, "gamma", seed=5)
    v = v.astype(np.float32)
    ptr = gen.csr_from_coo(x, rows)
    r0, r1 = sh.plan_row_shards_by_nnz(ptr, world)[rank]
    p, idx, val = sh.slice_csr(ptr, y, v, r0, r1)
    Q = np.stack([make_query(cols, 300 + i) for i in range(batch)])
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        eng = tks.SpMV(p, idx, val, r1 - r0, cols, k=k, device=rank, row_offset=r0, max_batch=max(batch, 1))
        sharded = tks.ShardedSpMV(eng, k, batch=batch, exchange=exchange)
        mode = sharded.exchange_mode
        out = []
        for rep in range(3):                      # three times: the exchange buffers (and parity slots) are reused
            eng.reset(Q if batch > 1 else Q[0])
            sharded.step(stream.cuda_stream)
            torch.cuda.synchronize()
            out = [eng.read_result(b) for b in range(batch)]
    q.put((rank, [(a.copy(), b.copy(), c) for a, b, c in out] + [mode]))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch,exchange,k", [(1, "auto", 100), (1, "auto", 300), (1, "auto", 1024), (1, "nccl", 100), (33, "auto", 100)])
def test_sharded_ranks_agree_with_global_topk(cuda_required, tks, orc, gen, batch, exchange, k):
    """exchange "auto" with one query per step = the peer-memory kernel (CUDA IPC windows over NVLink) when there are
    2 GPUs; "nccl" = all-gather + merge kernel; batched steps always take the all-gather path.  k = 300 makes the
    exchange kernel merge 600 keys (its bitonic branch), k = 1024 fills the windows (world * k = 2048)."""
    import torch
    import torch.multiprocessing as mp
    from conftest import make_query
    world = 2 if torch.cuda.device_count() >= 2 else 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q, exchange, k)) for r in range(world)]
    for p in procs:
        p.start()
    from conftest import collect_from_workers
    results = dict(collect_from_workers(q, procs, world, 300))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rows, cols = 20000, 1024
    x, y, v = gen.create_sparse_matrix(rows, cols, 20, "gamma", seed=5)
    v = v.astype(np.float32)
    modes = {results[r][-1] for r in range(world)}
    assert len(modes) == 1, "ranks disagree on the exchange mode"
    if world > 1 and batch == 1 and exchange == "auto":
        assert modes == {"peer"}, "peer-memory exchange was not set up on a multi-GPU box"
    else:
        assert modes == {"nccl"}
    for b in range(batch):
        yref = orc.spmv_f32(x, y, v, make_query(cols, 300 + b), rows)
        order = np.lexsort((np.arange(rows), -yref.astype(np.float64)))[:k]
        for rank in range(world):
            val, idx, cnt = results[rank][b]
            assert cnt == k
            # larger k: rows whose scores tie within fp32 rounding may swap across the k-th place
            assert len(set(idx.tolist()) ^ set(order.tolist())) <= (0 if k <= 100 else 4)
            np.testing.assert_allclose(val, yref[idx], rtol=1e-5)
            if batch > 1:      # the batched kernel's sums are sequential fp32: exact, in the exact order
                assert np.array_equal(idx, order.astype(np.uint32))
                assert np.array_equal(val.view(np.uint32), yref[order].view(np.uint32))


def _worker_fixed(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    from _pkg import pkg
    from conftest import make_query
    tks = pkg()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rows, cols = 30001, 1024
    x, y, v = tks.create_matrices.create_sparse_matrix(rows, cols, 20, "gamma", seed=6)
    s = tks.ShardedSpMVFixed(x, y, oracle.fx32_from_double(v), rows, cols, k=100, device=rank)
    out = []
    for qs in (1, 2):
        s.reset(oracle.query_fx32_from_f32(make_query(cols, qs)))
        s()
        val, idx = s.read_result()
        out.append((val.copy(), idx.copy(), s.idx_words.copy(), s.val_words.copy()))
    # pipelined form: six queries through submit / fetch with two in flight == the blocking verbs, words included
    qs = [oracle.query_fx32_from_f32(make_query(cols, 1 + (i % 2))) for i in range(6)]
    prev, got = None, []
    for v32 in qs:
        t = s.submit(v32)
        if prev is not None:
            val, idx = s.fetch(prev)
            got.append((val.copy(), idx.copy(), s.idx_words.copy(), s.val_words.copy()))
        prev = t
    val, idx = s.fetch(prev)
    got.append((val.copy(), idx.copy(), s.idx_words.copy(), s.val_words.copy()))
    for i, g in enumerate(got):
        ref = out[i % 2]
        assert all(np.array_equal(a, b) for a, b in zip(g, ref)), f"pipelined step {i} differs from the blocking verbs"
    q.put((rank, out))
    dist.barrier()
    s.close()
    dist.destroy_process_group()


def test_fixed_mode_partitions_over_ranks_equal_oracle(cuda_required, tks, orc, gen):
    """FPGA mode over NCCL: 32 partitions dealt out over the ranks (16 + 16 on 2 GPUs), result words all-gathered,
    the reference's merge on every rank: bit-exact against the oracle of the UNSHARDED matrix."""
    import torch
    import torch.multiprocessing as mp
    from conftest import make_query
    world = 2 if torch.cuda.device_count() >= 2 else 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fixed, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    from conftest import collect_from_workers
    results = dict(collect_from_workers(q, procs, world, 300))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rows, cols = 30001, 1024
    x, y, v = gen.create_sparse_matrix(rows, cols, 20, "gamma", seed=6)
    for i, qs in enumerate((1, 2)):
        o = orc.bscsr_topk(x, y, v, rows, make_query(cols, qs))
        for rank in range(world):
            val, idx, iw, vw = results[rank][i]
            assert np.array_equal(iw, o["idx_words"]) and np.array_equal(vw, o["val_words"])
            assert np.array_equal(idx, o["idx"][:100]) and np.array_equal(val, o["val"][:100])
