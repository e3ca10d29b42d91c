"""Per-step timing of the multi-GPU candidate exchange (run under torchrun, one rank per GPU):
local kernels (sample + main + select) and the exchange kernel are bracketed separately with CUDA events on
every rank; prints per-rank means and the distribution of the exchange time.  Measurement aid, not product code."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from _pkg import pkg  # noqa: E402


def main():
    tks = pkg()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows, cols, k, steps = 10_000_000, 1024, 100, 60
    eng = tks.SpMV(num_cols=cols, k=k, device=local)
    eng.generate_synthetic(rows, cols, 20, "gamma", seed=0, row_offset=rank * rows)
    sh = tks.ShardedSpMV(eng, k, batch=1, exchange="peer")
    rng = np.random.default_rng(1)
    q = rng.random((steps, cols)).astype(np.float32)
    dq = torch.from_numpy(q).cuda()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    s = stream.cuda_stream
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for mode in ("split", "fused"):
        torch.cuda.synchronize(); dist.barrier()
        for i in range(steps):
            eng.reset_device(dq[i].data_ptr(), 1, s)
            ev[i][0].record()
            if mode == "split":
                eng.run_async(k, s)
                ev[i][1].record()
                eng.peer_exchange_async(k, s)
            else:
                eng.run_exchange_async(k, s)
                ev[i][1].record()
            ev[i][2].record()
        torch.cuda.synchronize()
        a = np.array([ev[i][0].elapsed_time(ev[i][1]) for i in range(10, steps)])
        b = np.array([ev[i][1].elapsed_time(ev[i][2]) for i in range(10, steps)])
        t = np.array([ev[i][0].elapsed_time(ev[i + 1][0]) for i in range(10, steps - 1)])
        print(json.dumps({"rank": rank, "mode": mode, "local_ms_mean": a.mean(), "local_ms_std": a.std(),
                          "exchange_ms_mean": b.mean(), "exchange_ms_p10": float(np.percentile(b, 10)),
                          "exchange_ms_p90": float(np.percentile(b, 90)), "step_to_step_ms": t.mean()}), flush=True)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
