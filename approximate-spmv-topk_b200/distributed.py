"""Row-sharded Top-K SpMV over several GPUs of one box, one process per GPU.

The reference has no multi-device code; its only "scale-out" is the 32 row partitions of one FPGA whose
K-entry candidate lists the host merges (src/fpga/src/host_spmv_bscsr.cpp:399-448).  Here every rank owns a
contiguous row shard resident in its GPU's HBM and runs the fused kernel on it; the ONLY exchange per query
is an all-gather of the k best (score, row) candidates of every rank -- k 64-bit ordering keys, 800 bytes
for k = 100 -- followed by a merge kernel on every rank (top-k of a union = top-k of the per-shard top-k's).
torch.distributed (NCCL over NVLink/NVSwitch) is the plumbing; the keys never leave the device.

The same class runs under `gloo` with a host-side stand-in engine, which is how the CPU tests cover the
exchange and merge logic (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

import numpy as np

from . import sharding

KMAX = 1024   # stride (in keys) between the result lists of consecutive queries inside the engine


class _DevView:
    """Zero-copy torch view of engine-owned device memory."""

    def __init__(self, ptr, shape, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


_CUDA_STREAM_LEGACY = 1   # cudaStreamLegacy: the handle that NAMES the legacy default stream


def _current_stream_handle(torch):
    """torch's current stream as the C ABI wants it.  The ABI reads a NULL stream as "the handle's private stream", and
    torch's default stream IS handle 0: named by cudaStreamLegacy instead, the engine's kernels land on the stream the
    collectives and copies of the caller are ordered against."""
    h = torch.cuda.current_stream().cuda_stream
    return h if h else _CUDA_STREAM_LEGACY


class ShardedSpMV:
    """engine: a `spmv.SpMV` holding this rank's shard (uploaded or generated with row_offset = first row).

    step(k, stream): run on the shard, all-gather the candidates, merge; afterwards `engine.read_result(q)`
    returns the GLOBAL top-k on every rank."""

    def __init__(self, engine, k, batch=1, group=None, exchange="auto"):
        """exchange: "peer" = one kernel stores the candidates into every rank's CUDA-IPC-mapped window over NVLink and
        merges (tks_run_exchange_async: no NCCL launch, no separate merge launch); "nccl" = all-gather + merge kernel;
        "auto" = peer when it can be set up (NCCL backend, one query per step, world * k <= 2048), else nccl."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.engine, self.k, self.batch, self.group = engine, int(k), int(batch), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.gathered = None
        self.exchange_mode = "none" if exchange == "none" else "nccl"   # "none": timing experiments only (local top-k)
        want_peer = exchange in ("auto", "peer") and self.world > 1 and engine is not None and self.batch == 1 \
            and self.world <= 8 and self.world * self.k <= 2048 and dist.get_backend(group) == "nccl"
        if exchange == "peer" and not want_peer:
            raise ValueError("peer exchange needs the NCCL backend, batch == 1, world <= 8 and world * k <= 2048")
        if want_peer:
            ok, handles = 1, [None] * self.world
            try:
                mine = engine.peer_init(self.world, self.rank)
            except Exception as e:                     # e.g. IPC not permitted in this container
                ok, mine, self.peer_error = 0, b"", str(e)
            dist.all_gather_object(handles, (ok, mine), group=group)
            if all(h[0] for h in handles):
                try:
                    engine.peer_connect([h[1] for h in handles])
                except Exception as e:
                    ok, self.peer_error = 0, str(e)
            else:
                ok = 0
            flags = [None] * self.world
            dist.all_gather_object(flags, ok, group=group)   # every rank must agree on the mode
            if all(flags):
                self.exchange_mode = "peer"
            elif exchange == "peer":
                raise RuntimeError("peer exchange could not be set up: " + getattr(self, "peer_error", "a peer failed"))

    def _alloc(self, device):
        if self.gathered is None:
            self.gathered = self.torch.empty((self.world, self.batch, self.k), dtype=self.torch.int64, device=device)

    def exchange(self, mine):
        """mine: [batch, k] int64 keys of this rank -> [batch, world * k] keys regrouped by query."""
        self._alloc(mine.device)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered.view(-1), mine.reshape(-1), group=self.group)
        else:
            self.gathered[0].copy_(mine)
        if self.batch == 1:
            return self.gathered.view(1, self.world * self.k)           # already grouped by query
        return self.gathered.permute(1, 0, 2).reshape(self.batch, self.world * self.k).contiguous()

    def _stream(self, stream):
        """stream 0 would mean "the engine's private stream" to the C ABI, but NCCL orders its collectives against
        torch's CURRENT stream only: on the all-gather path the two must be the same stream."""
        if stream == 0 and self.world > 1 and self.exchange_mode == "nccl":
            return _current_stream_handle(self.torch)
        return stream

    def submit(self, dptr, stream=0, query_ready=True):
        """Pipelined step (tks_submit): the query at device pointer `dptr` is enqueued and consecutive steps overlap;
        with several ranks the select kernel of every step also exchanges and merges (peer mode).  The all-gather
        path has no pipelined form and runs the step in stream order."""
        eng, k = self.engine, self.k
        if self.batch != 1:
            raise ValueError("pipelined submits take one query per step")
        if self.exchange_mode == "peer" or self.world == 1 or self.exchange_mode == "none":
            eng.submit(dptr, k, stream, exchange=self.exchange_mode == "peer", query_ready=query_ready)
            self.pipelined = True
            return
        stream = self._stream(stream)
        eng.reset_device(dptr, 1, stream)
        self.step(stream)

    def submit_host(self, vec):
        """Pipelined step fed from host memory (tks_submit_host): returns a ticket for fetch().  Peer mode or one rank."""
        if self.batch != 1 or not (self.exchange_mode == "peer" or self.world == 1 or self.exchange_mode == "none"):
            raise ValueError("host-fed pipelined steps need one query per step and the peer-memory exchange (or one rank)")
        self.pipelined = True
        return self.engine.submit_host(vec, self.k, exchange=self.exchange_mode == "peer")

    def fetch(self, ticket):
        """(values, GLOBAL row indices, count) of the step submit_host() gave `ticket` for; identical on every rank."""
        return self.engine.fetch(ticket)

    def wait(self, stream=0):
        """Make `stream` wait for the last submitted step's (global) result."""
        if getattr(self, "pipelined", False):
            self.engine.pipeline_wait(stream)

    def step(self, stream=0):
        eng, k, B = self.engine, self.k, self.batch
        stream = self._stream(stream)
        if self.exchange_mode == "peer":
            eng.run_exchange_async(k, stream)
            return
        eng.run_async(k, stream)
        if self.world == 1 or self.exchange_mode == "none":
            return
        kp, _ = eng.result_keys_device(0)
        mine = self.torch.as_tensor(_DevView(kp, (B, KMAX)), device="cuda")[:, :k]
        if B > 1:
            mine = mine.contiguous()
        regrouped = self.exchange(mine)
        self._keep = regrouped                                           # alive until the merge has run
        if B == 1:
            eng.merge_keys_device(regrouped.data_ptr(), self.world * k, k, 0, stream)
        else:
            eng.merge_keys_batched_device(regrouped.data_ptr(), self.world * k, B, k, stream)


def plan_partition_shards(num_rows, partitions, world):
    """FPGA mode: the P row partitions of host_spmv_bscsr.cpp:136-141 (rows_per_part = ceil(N / P)) are dealt
    out contiguously, P / world per rank.  Returns (rows_per_part, parts_per_rank, [(row_begin, row_end)] per rank);
    every shard is handed to its engine with P / world partitions and parts_per_rank * rows_per_part (virtual) rows, so
    the partition boundaries are the ones of the unsharded matrix."""
    if partitions % world != 0:
        raise ValueError(f"partitions={partitions} must be a multiple of the number of ranks ({world})")
    rpp = (int(num_rows) + partitions - 1) // partitions
    ppr = partitions // world
    return rpp, ppr, [(min(r * ppr * rpp, int(num_rows)), min((r + 1) * ppr * rpp, int(num_rows))) for r in range(world)]


class ShardedSpMVFixed:
    """FPGA-semantics engine over several GPUs, one process per GPU: rank r owns partitions [r*P/N, (r+1)*P/N) of the
    reference's 32 (the rows [r0, r1) of plan_partition_shards), runs the unchanged kernels on them, and the ONLY
    exchange per query is an all-gather of the result words (P/N x local_k x 128 bytes per rank: 32 KB in total for the
    reference's knobs) followed by the reference's host merge (host_spmv_bscsr.cpp:399-448) over all P partitions on
    every rank -- bit-identical to one device holding all partitions.

    x, y, val32: the WHOLE row-sorted COO (every rank slices its own part); engine_factory builds the local
    `spmv.SpMVFixed` (tests inject a stand-in under gloo)."""

    def __init__(self, x, y, val32, num_rows, num_cols, k=100, fixed_width=20, partitions=32, local_k=8,
                 limited_finished_rows=4, drift_free=False, device=0, device_pack=True, group=None, engine_factory=None):
        import torch
        import torch.distributed as dist
        from . import capi, spmv
        self.torch, self.dist, self.capi, self.group = torch, dist, capi, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.k, self.P, self.Kp = int(k), int(partitions), int(local_k)
        self.B = capi.bscsr_packet_size(fixed_width)
        self.rpp, self.ppr, shards = plan_partition_shards(num_rows, partitions, self.world)
        self.r0, self.r1 = shards[self.rank]
        x = np.asarray(x)
        b, e = np.searchsorted(x, self.r0, side="left"), np.searchsorted(x, self.r1, side="left")
        lx = (x[b:e] - self.r0).astype(np.uint32)
        factory = engine_factory or (lambda *a, **kw: spmv.SpMVFixed(*a, **kw))
        self.engine = factory(lx, np.asarray(y)[b:e], np.asarray(val32)[b:e], self.ppr * self.rpp, num_cols, k=k,
                              fixed_width=fixed_width, partitions=self.ppr, local_k=local_k,
                              limited_finished_rows=limited_finished_rows, drift_free=drift_free, device=device,
                              device_pack=device_pack)
        self.nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
        # global first row of every partition: local first_row (host:145) + the shard's row offset, gathered once
        mine = (np.asarray(self.engine.first_row_array(), np.int64) + self.r0).astype(np.int32)
        self.first_row = self._all_gather(mine.view(np.uint32)).reshape(self.P)

    def _all_gather(self, words):
        """uint32 array of equal size on every rank -> concatenation over ranks (rank order)."""
        t = self.torch.from_numpy(np.ascontiguousarray(words).view(np.int32).reshape(-1).copy())
        if self.world == 1:
            return t.numpy().view(np.uint32)
        if self.nccl:
            t = t.cuda()
        out = self.torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().numpy().view(np.uint32)

    def reset(self, vec32):
        return self.engine.reset(vec32)

    def __call__(self, debug=0):
        """operator(): run the local partitions, exchange the result words, merge.  Returns the local kernel ns."""
        if self.nccl and self.world > 1:
            return self._call_device_gather()
        ns = self.engine()
        iw, vw = self.engine.read_partition_results()
        allw = self._all_gather(np.concatenate([iw.reshape(-1), vw.reshape(-1)]))
        per = self.ppr * self.Kp * 16
        allw = allw.reshape(self.world, 2, per)
        self.idx_words = allw[:, 0, :].reshape(self.P, self.Kp, 16)
        self.val_words = allw[:, 1, :].reshape(self.P, self.Kp, 16)
        self._val, self._idx = self.capi.merge_partition_words(self.idx_words, self.val_words, self.first_row, self.B, self.k)
        return ns

    def _call_device_gather(self):
        """NCCL ranks: the kernels, the all-gather of the result words (device to device, straight from the engine's result
        block) and ONE device-to-host copy of all P partitions' words are enqueued on torch's current stream; the host
        waits once, then merges."""
        torch = self.torch
        stream = torch.cuda.current_stream()
        if getattr(self, "_gath", None) is None:
            ptr, n = self.engine.partition_words_device()
            self._words = torch.as_tensor(_DevView(ptr, (n,), "<i4"), device="cuda")
            self._gath = torch.empty(self.world * n, dtype=torch.int32, device="cuda")
            self._host = torch.empty(self.world * n, dtype=torch.int32).pin_memory()
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._ev[0].record(stream)
        self.engine.run_async(self.k, _current_stream_handle(torch))
        self._ev[1].record(stream)
        self.dist.all_gather_into_tensor(self._gath, self._words, group=self.group)
        self._host.copy_(self._gath, non_blocking=True)
        stream.synchronize()
        per = self.ppr * self.Kp * 16
        allw = self._host.numpy().view(np.uint32).reshape(self.world, 2, per)
        self.idx_words = allw[:, 0, :].reshape(self.P, self.Kp, 16)
        self.val_words = allw[:, 1, :].reshape(self.P, self.Kp, 16)
        self._val, self._idx = self.capi.merge_partition_words(self.idx_words, self.val_words, self.first_row, self.B, self.k)
        return self._ev[0].elapsed_time(self._ev[1]) * 1e6

    # ---- pipelined form: two queries in flight -------------------------------------------------------------------
    def submit(self, vec32):
        """reset(vec) + operator() without waiting: the query's copy, transform and sample run beside the previous
        query's stream and replay kernels (tks_submit), the all-gather of the result words and ONE device-to-host copy
        follow on torch's current stream, and the host returns at once with a ticket for fetch().  Two queries are kept
        in flight: fetch(t) before submit() number t + 2.  Without NCCL (gloo tests, one rank without CUDA) the call
        runs the blocking verbs and fetch() returns what they produced."""
        self._n_sub = getattr(self, "_n_sub", 0) + 1
        ticket = self._n_sub
        slot = ticket % 2
        if not hasattr(self, "_kept"):
            self._kept = [None, None]
        if not (self.nccl or (self.world == 1 and self.torch.cuda.is_available() and hasattr(self.engine, "submit"))):
            self.reset(vec32)
            self()
            self._kept[slot] = (ticket, None, (self._val, self._idx, self.idx_words, self.val_words))
            return ticket
        torch = self.torch
        stream = torch.cuda.current_stream()
        if getattr(self, "_pslots", None) is None:
            ptr, n = self.engine.partition_words_device()
            self._pwords = torch.as_tensor(_DevView(ptr, (n,), "<i4"), device="cuda")
            cols = int(self.engine.num_cols)
            self._pslots = [dict(hq=torch.empty(cols, dtype=torch.int32).pin_memory(),
                                 dq=torch.empty(cols, dtype=torch.int32, device="cuda"),
                                 gath=torch.empty(self.world * n, dtype=torch.int32, device="cuda"),
                                 host=torch.empty(self.world * n, dtype=torch.int32).pin_memory(),
                                 done=torch.cuda.Event()) for _ in range(2)]
        sl = self._pslots[slot]
        if self._kept[slot] is not None:
            sl["done"].synchronize()   # the slot's query of two submits ago was never fetched: its buffers are reused now
        sl["hq"].numpy()[:] = np.ascontiguousarray(vec32, np.uint32).view(np.int32).reshape(-1)
        sl["dq"].copy_(sl["hq"], non_blocking=True)
        self.engine.submit(sl["dq"].data_ptr(), self.k, _current_stream_handle(torch))
        if self.world > 1:
            self.dist.all_gather_into_tensor(sl["gath"], self._pwords, group=self.group)
            sl["host"].copy_(sl["gath"], non_blocking=True)
        else:
            sl["host"].copy_(self._pwords, non_blocking=True)
        sl["done"].record(stream)
        self._kept[slot] = (ticket, sl, None)
        return ticket

    def fetch(self, ticket):
        """read_result() of submit()'s `ticket`: (raw values uint32[n], GLOBAL row indices uint32[n]); also sets
        idx_words / val_words like operator()."""
        kept = getattr(self, "_kept", [None, None])[int(ticket) % 2]
        if kept is None or kept[0] != ticket:
            raise self.capi.TksError(self.capi.TKS_ESTATE, f"the result of ticket {ticket} is gone: two queries are kept")
        self._kept[int(ticket) % 2] = None
        _, sl, ready = kept
        if ready is not None:
            self._val, self._idx, self.idx_words, self.val_words = ready
            return self._val, self._idx
        sl["done"].synchronize()
        per = self.ppr * self.Kp * 16
        allw = sl["host"].numpy().view(np.uint32).reshape(self.world, 2, per).copy()
        self.idx_words = allw[:, 0, :].reshape(self.P, self.Kp, 16)
        self.val_words = allw[:, 1, :].reshape(self.P, self.Kp, 16)
        self._val, self._idx = self.capi.merge_partition_words(self.idx_words, self.val_words, self.first_row, self.B, self.k)
        return self._val, self._idx

    def read_result(self):
        """(raw values uint32[n], GLOBAL row indices uint32[n]), n <= k; identical on every rank."""
        return self._val, self._idx

    def read_partition_results(self):
        return self.idx_words, self.val_words

    def close(self):
        self.engine.close()


def merge_gathered_host(gathered, k):
    """Host stand-in of the merge kernel for the gloo tests: gathered [batch, world*k] uint64 keys."""
    g = np.asarray(gathered).view(np.uint64)
    return np.stack([np.pad(m, (0, k - m.size)) for m in (sharding.merge_topk_host([row], k) for row in g)])
