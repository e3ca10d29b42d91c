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


class ShardedSpMV:
    """engine: a `spmv.SpMV` holding this rank's shard (uploaded or generated with row_offset = first row).

    step(k, stream): run on the shard, all-gather the candidates, merge; afterwards `engine.read_result(q)`
    returns the GLOBAL top-k on every rank."""

    def __init__(self, engine, k, batch=1, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.engine, self.k, self.batch, self.group = engine, int(k), int(batch), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.gathered = None

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

    def step(self, stream=0):
        eng, k, B = self.engine, self.k, self.batch
        eng.run_async(k, stream)
        if self.world == 1:
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


def merge_gathered_host(gathered, k):
    """Host stand-in of the merge kernel for the gloo tests: gathered [batch, world*k] uint64 keys."""
    g = np.asarray(gathered).view(np.uint64)
    return np.stack([np.pad(m, (0, k - m.size)) for m in (sharding.merge_topk_host([row], k) for row in g)])
