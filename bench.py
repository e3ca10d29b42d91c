#!/usr/bin/env python
"""bench.py -- fused Top-K SpMV throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

A "step" = one Top-K SpMV query over the resident matrix:
  N = 1   workload cfg2: synthetic 10M x 1024, gamma ~20 nnz/row, fp32 CSR, single query, k = 100
  N > 1   the same workload weak-scaled: every rank holds a 10M-row shard of an (N x 10M) x 1024 matrix of the
          same law (per-GPU work fixed, "scaling": "weak"), K candidates per rank all-gathered (NCCL) and
          merged on every rank -- so that the per-N values of one command are comparable
  --workload cfg4: synthetic 200M x 1024, uniform ~40 nnz/row, fp32, k = 100, rows sharded evenly over the
          N ranks (fixed total work, "scaling": "strong"); fits one GPU too (64 GB)
  --workload cfg3: the cfg2 matrix as 20-bit BS-CSR packets, 32 partitions x local K=8 (FPGA semantics)
`value` = non-zeros processed per second by the whole job with the matrix and the queries resident in HBM;
`e2e`   = the same through the reference-facing calls reset(vec) / operator() / read_result with HOST
          buffers (query H2D and result D2H inside the timed region).
The matrix is generated in HBM (tks_generate_synthetic: the law of the reference's create_matrices.py);
it is far larger than the 126 MB L2, so no flush is needed between steps.

--impl reference times the reference's CPU path on the box's host cores (see cpu_reference()).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "cfg2": dict(rows=10_000_000, cols=1024, deg=20, dist="gamma", mode="float",
                 name="cfg2: synthetic 10M x 1024, gamma ~20 nnz/row, fp32 CSR, single query, k=100"),
    "cfg2h": dict(rows=10_000_000, cols=1024, deg=20, dist="gamma", mode="float", half=True,
                  name="cfg2h: cfg2 with half-precision matrix values and query (the reference's -a GPU mode), fp32 accumulation, k=100"),
    "cfg2b": dict(rows=10_000_000, cols=1024, deg=20, dist="gamma", mode="float", bf16=True,
                  name="cfg2b: cfg2 with bfloat16 matrix values and query, fp32 accumulation, k=100"),
    "cfg3": dict(rows=10_000_000, cols=1024, deg=20, dist="gamma", mode="fixed",
                 name="cfg3: cfg2 matrix in 20-bit BS-CSR packets, 32 partitions x local K=8 (FPGA semantics)"),
    "cfg4": dict(rows=200_000_000, cols=1024, deg=40, dist="uniform", mode="float",
                 name="cfg4: synthetic 200M x 1024, uniform ~40 nnz/row, fp32, k=100, row-sharded + K-candidate allgather"),
    "cfg5": dict(rows=50_000_000, cols=1024, deg=20, dist="gamma", mode="batched", batch=64,
                 name="cfg5: batched 64 queries x 50M x 1024 gamma ~20 nnz/row (one matrix read per 32 queries), fp32, k=100"),
}
K = 100
SEED = 0


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            j = json.loads(p.read_text())
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  The timed region of a
    default run is a few milliseconds, far below nvidia-smi's sampling period, so the sampler polls NVML itself
    (nvidia_ml_py) from a thread about once per millisecond; mark(name) labels what follows, and the summary reports
    the samples that fell inside the label "timed" next to those of the whole measurement."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake_slowdown", 0x80), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.label, self.stop_flag, self.thread, self.err = index, [], "setup", False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device ordinal follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.nv, self.dev = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception as e:   # no NVML: say so in the line instead of failing the bench
            self.nv, self.err = None, repr(e)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                self.rows.append((self.label, float(mhz), int(why)))
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(0.0007)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def mark(self, label):
        self.label = label

    def sample_now(self, label):
        """One sample taken by the calling thread (the polling thread may not get a turn inside a 4 ms region)."""
        if self.nv is None:
            return
        try:
            nv = self.nv
            mhz = nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)
            try:
                why = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
            except Exception:
                why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
            self.rows.append((label, float(mhz), int(why)))
        except Exception as e:
            self.err = repr(e)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["NVML unavailable: " + str(self.err)]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        timed = [r for r in self.rows if r[0] == "timed"]

        def summary(rows):
            why = 0
            for r in rows:
                why |= r[2]
            return {"sm_mhz": statistics.median([r[1] for r in rows]) if rows else None, "samples": len(rows),
                    "reasons": sorted(n for n, bit in self.REASONS if why & bit)}
        t, a = summary(timed), summary(self.rows)
        return {"sm_mhz": t["sm_mhz"] if t["samples"] else a["sm_mhz"], "sm_max_mhz": self.max_mhz,
                "samples": t["samples"], "reasons": sorted(set(t["reasons"]) | set(a["reasons"])),
                "whole_measurement": a, "source": "NVML polled every ~1 ms; `samples` = inside the timed region"}


def make_queries(cols, n, seed0=1):
    """test_cpu.py:99-100: U[0,1)^C / L2 norm; seeds 1..n (the reference is unseeded)."""
    out = np.zeros((n, cols), np.float32)
    for i in range(n):
        rng = np.random.default_rng(seed0 + i)
        v = rng.random(cols)
        out[i] = (v / np.linalg.norm(v)).astype(np.float32)
    return out


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU implementation on the host cores
# ---------------------------------------------------------------------------------------------

def cpu_reference(ptr, idx, val, queries, k, max_seconds=25.0, min_steps=1):
    """Times the reference's own CPU Top-K SpMV on this box.

    The reference's CPU path is sparse_dot_topn.awesome_cossim_topn(..., use_threads=True, n_jobs=40)
    (test_cpu.py:104); that dependency is absent and un-pinned, so the arm runs the reference's other CPU
    implementation of the same path, spmv_coo_gold_top_k + sort_tuples (gold_algorithms.hpp:188-246,
    evaluation_utils.hpp:40-62): the reference's OWN code from oracle/_ref/libref_gold.so when it was
    built (kind "reference"), else the oracle's restatement of it (kind "port").  Like the threaded
    sparse_dot_topn, rows are split into one contiguous block per host thread, each block is reduced
    with the reference routine, and the per-block top-k lists are merged.
    Returns (seconds_per_query_list, kind, cores, results)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    rows = ptr.size - 1
    cores = os.cpu_count() or 1
    use_ref = oracle.ref_gold() is not None
    kind = "reference" if use_ref else "port"
    fn = oracle.ref_gold_topk_f32 if use_ref else oracle.gold_topk_f32
    deg = np.diff(ptr.astype(np.int64))
    x = np.repeat(np.arange(rows, dtype=np.uint32), deg)
    cuts = [int(ptr[rows * t // cores]) for t in range(cores + 1)]
    blocks = [(cuts[t], cuts[t + 1]) for t in range(cores) if cuts[t + 1] > cuts[t]]
    times, results = [], []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        t_start = time.perf_counter()
        for q in queries:
            t0 = time.perf_counter()
            parts = list(ex.map(lambda b: fn(x[b[0]:b[1]], idx[b[0]:b[1]], val[b[0]:b[1]], q, k), blocks))
            ai = np.concatenate([p[0] for p in parts]); av = np.concatenate([p[1] for p in parts])
            order = np.lexsort((-ai.astype(np.int64), -av.astype(np.float64)))[:k]
            results.append((ai[order], av[order]))
            times.append(time.perf_counter() - t0)
            if len(times) >= min_steps and time.perf_counter() - t_start > max_seconds:
                break
    return times, kind, cores, results


def standin_f64(ptr, idx, val, queries, k, max_seconds=15.0):
    """BASELINE.md section 3 item 2 / test_cpu.py:91-105 semantics with the absent sparse_dot_topn replaced by what that
    call computes for a one-column right-hand side: float64 scipy `csr @ vec`, entries <= 0 dropped, global top-k by
    argpartition (test_cpu.py's own stand-in, imported from the repo-root driver).  The CSR is built before the clock
    starts, as test_cpu.py builds it before its timed loop.  scipy's product is single-threaded."""
    import scipy.sparse as sp
    import test_cpu
    a = sp.csr_matrix((val.astype(np.float64), idx.astype(np.int64), ptr.astype(np.int64)),
                      shape=(ptr.size - 1, int(queries[0].size)))
    times, t_start = [], time.perf_counter()
    for q in queries:
        t0 = time.perf_counter()
        test_cpu.topk_standin(a, q.astype(np.float64), k)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > max_seconds:
            break
    return times


def reference_matrix(wl, rows_wanted, args):
    """The benchmark matrix for the CPU arm: generated in HBM by the same generator the GPU arm uses (input
    preparation, nothing of it is timed) and copied to the host; without a GPU, the NumPy restatement of the
    reference's generator on a bounded number of rows.  Returns (ptr, idx, val, rows, how)."""
    from _pkg import pkg
    tks = pkg()
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    rows = rows_wanted
    while rows > 1_000_000 and rows * wl["deg"] * 24 > avail // 2:      # ptr + idx + val + COO rows + slack
        rows //= 2
    try:
        import torch
        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    if gpu:
        eng = tks.SpMV(num_cols=wl["cols"], k=K)
        eng.generate_synthetic(rows, wl["cols"], wl["deg"], wl["dist"], seed=SEED)
        ptr, idx, val = eng.download_csr()
        eng.close()
        return ptr, idx, val, rows, "generated in HBM by tks_generate_synthetic (the GPU arm's matrix, seed 0) and copied to the host"
    rows = min(rows, args.ref_rows or 2_000_000)
    x, y, v = tks.create_matrices.create_sparse_matrix(rows, wl["cols"], wl["deg"], wl["dist"], seed=SEED)
    return tks.create_matrices.csr_from_coo(x, rows), y, v.astype(np.float32), rows, "NumPy restatement of create_matrices.py (no GPU here)"


def reference_arm(args):
    """`--impl reference`: rank 0 only.  The reference's CPU Top-K SpMV over the GPU arm's configuration -- the whole
    matrix (N = 1: 10M rows; N > 1: the N x 10M rows of the weak-scaled run), every step one query -- on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl_key = args.workload or "cfg2"
    wl = WORKLOADS[wl_key]
    world = max(1, args.gpus)
    weak = wl_key in ("cfg2", "cfg2h", "cfg2b")
    rows_cfg = (args.rows or wl["rows"]) * (world if weak else 1)
    rows_wanted = min(rows_cfg, args.ref_rows) if args.ref_rows else rows_cfg
    ptr, idx, val, rows, how = reference_matrix(wl, rows_wanted, args)
    nnz = int(ptr[-1])
    queries = make_queries(wl["cols"], args.warmup + args.steps)
    cpu_reference(ptr, idx, val, queries[:args.warmup], K, max_seconds=1e9)
    times, kind, cores, _ = cpu_reference(ptr, idx, val, queries[args.warmup:], K, max_seconds=600.0, min_steps=args.steps)
    sec = sum(times) / len(times)
    value = nnz / sec
    stand = standin_f64(ptr, idx, val, queries[args.warmup:args.warmup + 5], K) if rows <= 20_000_000 else []
    sample = (f"{rows} of {rows_cfg} rows of the configuration ({nnz} nnz; {how}), {len(times)} queries, "
              f"reference spmv_coo_gold_top_k over {cores} row blocks in {cores} host threads "
              f"(sparse_dot_topn is absent from the image)")
    cpu = {"value": value, "unit": "nnz/s", "cores": cores, "kind": kind, "sample": sample, "same_config": rows == rows_cfg}
    if stand:
        ssec = sum(stand) / len(stand)
        cpu["stand_in"] = {"value": nnz / ssec, "unit": "nnz/s", "cores": 1, "kind": "port", "ms_per_query": ssec * 1e3,
                           "what": "test_cpu.py:91-105 with sparse_dot_topn (absent, un-pinned) replaced by float64 scipy csr @ vec + "
                                   "argpartition top-k (BASELINE.md 3.2), same matrix, %d queries; scipy's product runs on one thread" % len(stand)}
    line = {"impl": "reference", "metric": "topk_spmv_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": wl["name"] + (f", weak-scaled: {world} shards of {wl['rows']} rows" if weak and world > 1 else ""),
                                            "rows": rows, "cols": wl["cols"], "nnz": nnz, "k": K},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

def ours(args):
    import torch
    import torch.distributed as dist
    from _pkg import pkg
    tks = pkg()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl_key = args.workload or "cfg2"
    wl = WORKLOADS[wl_key]
    weak = wl_key in ("cfg2", "cfg2h", "cfg2b")                          # cfg2: 10M rows PER RANK; cfg4 / cfg5: the stated total, sharded
    rows_total = (args.rows or wl["rows"]) * (world if weak else 1)
    cols = wl["cols"]
    shards = tks.sharding.plan_row_shards_even(rows_total, world)
    r0, r1 = shards[rank]
    peak_gbs, peak_src = measured_peaks()
    # a non-default torch stream: everything of a step (query copy, our kernels, NCCL, merge) is enqueued on it
    # and the CUDA events that time the region are recorded on it (a NULL stream would mean "the handle's own")
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    nsteps = args.warmup + args.steps
    queries = make_queries(cols, nsteps)

    if wl["mode"] == "fixed" and world > 1:
        line = fixed_sharded_workload(tks, torch, dist, args, world, rank, local, args.steps, args.warmup)
        if rank == 0:
            print(json.dumps(line), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    if wl["mode"] == "fixed":
        return ours_fixed(args, tks, wl, rows_total, queries, peak_gbs, peak_src)
    if wl["mode"] == "batched":
        return ours_batched(args, tks, wl, rows_total, peak_gbs, peak_src, world, rank, local, stream)

    line = float_workload(tks, torch, dist, wl_key, args, world, rank, local, tstream, peak_gbs, peak_src,
                          steps=args.steps, warmup=args.warmup, full=True)
    if world == 1 and wl_key == "cfg2" and not args.no_sub:
        # the default single-GPU line also carries BASELINE configs 3 and 5 (reduced legs: value, e2e, roofline)
        wl3 = WORKLOADS["cfg3"]
        sub3 = fixed_workload(args, tks, wl3, wl3["rows"], make_queries(wl3["cols"], 3 + min(args.steps, 10)), peak_gbs, peak_src,
                              min(args.steps, 10), 3, False)
        torch.cuda.set_stream(tstream)
        wl5 = WORKLOADS["cfg5"]
        sub5 = batched_workload(tks, torch, dist, args, wl5, wl5["rows"], peak_gbs, peak_src, world, rank, local, stream,
                                steps=min(args.steps, 5), warmup=3, full=False)
        line["cfg3"], line["cfg5"] = sub3, sub5
    if world > 1 and wl_key == "cfg2" and not args.no_cfg4:
        # BASELINE config 4 rides along in every multi-GPU line of the default run: 200M x 1024 uniform-40, rows / N
        sub = float_workload(tks, torch, dist, "cfg4", args, world, rank, local, tstream, peak_gbs, peak_src,
                             steps=min(args.steps, 10), warmup=3, full=False)
        if rank == 0:
            line["cfg4"] = sub
        # ... and BASELINE config 5: 64 queries x 50M x 1024 gamma, rows / N, one matrix pass per 32 queries
        wl5 = WORKLOADS["cfg5"]
        sub5 = batched_workload(tks, torch, dist, args, wl5, wl5["rows"], peak_gbs, peak_src, world, rank, local, stream,
                                steps=min(args.steps, 5), warmup=3, full=False)
        if rank == 0:
            line["cfg5"] = sub5
        # ... and BASELINE config 3 with its 32 partitions dealt out over the ranks
        if 32 % world == 0:
            sub3 = fixed_sharded_workload(tks, torch, dist, args, world, rank, local, min(args.steps, 10), 3)
            if rank == 0:
                line["cfg3"] = sub3
    if rank == 0:
        print(json.dumps(line), flush=True)
    recs = [line or {}] + [(line or {}).get(k) or {} for k in ("cfg3", "cfg4", "cfg5")]
    ok = all(r.get("parity_n", True) is not False for r in recs)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not ok:
        raise SystemExit("parity check failed (see parity_n / parity in the line above)")


def aligned_start(torch, dist, world):
    """The ranks leave a barrier hundreds of microseconds apart (host scheduling), and with a per-step exchange the
    first step of the timed region would wait for the last rank to arrive -- a start-up skew charged to K = 20 short
    steps.  All ranks are processes of one host, so they agree on a wall-clock instant a few milliseconds ahead and
    spin until it: every rank's timed region then opens within microseconds of the others'."""
    if world <= 1:
        return
    t = torch.tensor([time.time() + 0.004], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    torch.cuda.synchronize()
    t_go = float(t.item())
    while time.time() < t_go:
        pass


def merge_lists(lists, k):
    """Per-rank (rows, scores) lists -> global top-k under (score desc, row asc)."""
    ai = np.concatenate([np.asarray(a[0], np.int64) for a in lists])
    av = np.concatenate([np.asarray(a[1], np.float32) for a in lists])
    order = np.lexsort((ai, -av.astype(np.float64)))[:k]
    return ai[order], av[order]


def lists_agree(idx, val, gi, gv, rtol=1e-5):
    """The north-star bar: scores within rtol, index sets identical except where scores tie within rtol of the k-th."""
    idx, gi = np.asarray(idx, np.int64), np.asarray(gi, np.int64)
    val, gv = np.asarray(val, np.float32), np.asarray(gv, np.float32)
    if idx.size != gi.size:
        return False, {"reason": f"{idx.size} results vs {gi.size} expected"}
    kth = float(gv[-1])
    tol = abs(kth) * rtol + 1e-7
    only_e = [(int(i), float(v)) for i, v in zip(idx, val) if i not in set(gi.tolist())]
    only_g = [(int(i), float(v)) for i, v in zip(gi, gv) if i not in set(idx.tolist())]
    ties_ok = all(abs(v - kth) <= tol for _, v in only_e + only_g)
    scores_ok = bool(np.allclose(np.sort(val)[::-1], np.sort(gv)[::-1], rtol=rtol, atol=1e-7))
    return bool(ties_ok and scores_ok), {"set_difference": len(only_e) + len(only_g), "k_th_score": kth,
                                         "max_abs_score_diff": float(np.max(np.abs(np.sort(val)[::-1] - np.sort(gv)[::-1])))}


def float_workload(tks, torch, dist, wl_key, args, world, rank, local, tstream, peak_gbs, peak_src, steps, warmup, full):
    """One float workload (cfg2 / cfg2h / cfg2b weak-scaled, cfg4 strong-scaled) on `world` ranks.  Returns the record
    (rank 0) or None.  full = the main line (e2e, roofline leg, CPU baseline); otherwise a sub-record with value,
    per-step statistics and parity."""
    wl = WORKLOADS[wl_key]
    weak = wl_key in ("cfg2", "cfg2h", "cfg2b")                          # cfg2: 10M rows PER RANK; cfg4: the stated total, sharded
    rows_total = (args.rows or wl["rows"]) * (world if weak else 1)
    cols = wl["cols"]
    r0, r1 = tks.sharding.plan_row_shards_even(rows_total, world)[rank]
    stream = tstream.cuda_stream
    nsteps = warmup + steps
    queries = make_queries(cols, nsteps)
    half = bool(wl.get("half", False) or wl.get("bf16", False))      # 16-bit values
    eng = tks.SpMV(num_cols=cols, k=K, device=local, half=bool(wl.get("half", False)), bf16=bool(wl.get("bf16", False)))
    t0 = time.perf_counter()
    eng.generate_synthetic(r1 - r0, cols, wl["deg"], wl["dist"], seed=SEED, row_offset=r0)
    gen_s = time.perf_counter() - t0
    st = eng.stats()
    nnz_local = int(st.nnz)
    nnz_t = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz_total = int(nnz_t.item())

    dq = torch.from_numpy(queries).cuda()            # queries resident in HBM for `value`
    torch.cuda.synchronize()
    # run + exchange of the K candidates + merge (nothing to exchange at N=1): the select kernel itself stores the
    # candidates into every rank's IPC window over NVLink, waits and merges when the ranks can map each other's windows,
    # else NCCL all-gather + merge kernel; TKS_EXCHANGE=nccl forces that
    sharded = tks.ShardedSpMV(eng, K, batch=1, exchange=os.environ.get("TKS_EXCHANGE", "auto"))
    pipelined = os.environ.get("TKS_BENCH_PIPELINE", "1") != "0" and (world == 1 or sharded.exchange_mode in ("peer", "none"))

    def step(i):
        if pipelined:
            # consecutive queries overlap (tks_submit): sample(i+1) beside main(i), select(i) beside main(i+1)
            sharded.submit(dq[i].data_ptr(), stream, query_ready=True)
        else:
            eng.reset_device(dq[i].data_ptr(), 1, stream)
            sharded.step(stream)

    for i in range(warmup):
        step(i)
    sharded.wait(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    aligned_start(torch, dist, world)
    sampler.mark("timed")
    e0.record()
    for i in range(steps):
        step(warmup + i)
    sharded.wait(stream)                             # the last query's select (+ exchange + merge) is inside the region
    e1.record()
    if rank == 0:
        sampler.sample_now("timed")                  # the host runs ahead: the device is still inside the timed steps here
    torch.cuda.synchronize()
    sampler.mark("after")
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / steps
    value = nnz_total / (ms_step * 1e-3)
    val_last, idx_last, cnt_last = eng.read_result()
    per_step = None
    if pipelined:
        # completion time of every step's select kernel (%globaltimer, written by the kernel): intervals = per-step times
        stamps = eng.pipeline_stamps(steps).astype(np.int64)
        if stamps.shape[0] >= 3:
            d = np.diff(stamps[:, 6]) * 1e-6
            us = lambda a: round(float(np.mean(a)) * 1e-3, 2)
            per_step = {"mean_ms": float(d.mean()), "std_ms": float(d.std()), "min_ms": float(d.min()), "max_ms": float(d.max()),
                        "n": int(d.size), "source": "intervals between the select kernels' %globaltimer stamps (this rank)",
                        # where the kernels of a step sit relative to each other (mean over the timed steps, microseconds)
                        "timeline_us": {"sample": us(stamps[:, 1] - stamps[:, 0]), "main": us(stamps[:, 3] - stamps[:, 2]),
                                        "select_after_main_end": us(stamps[:, 6] - stamps[:, 3]),
                                        "select": us(stamps[:, 6] - stamps[:, 5]),
                                        "sample_end_before_main_begin": us(stamps[:, 2] - stamps[:, 1]),
                                        "main_begin_after_previous_main_end": us(stamps[1:, 2] - stamps[:-1, 3])}}

    # ---- parity at N ranks: every rank runs the reference's CPU gold over its own shard (or a bounded sample of it)
    q_chk = queries[nsteps - 1]
    if wl.get("half"):
        q_chk = q_chk.astype(np.float16).astype(np.float32)        # the engine rounds the query to the storage type
    elif wl.get("bf16"):
        b = q_chk.view(np.uint32).astype(np.uint64)
        q_chk = (((b + 0x7FFF + ((b >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(np.float32)
    parity = parity_check(tks, torch, dist, eng, sharded, q_chk, queries[nsteps - 1], idx_last, val_last, world, rank,
                          r0, r1, stream, whole_shard=(r1 - r0) <= 12_000_000)

    stats = eng.stats()
    alg_bytes_local = int(stats.algorithmic_bytes)
    hq = [np.ascontiguousarray(q) for q in queries]
    rec = None
    e2e_ms_step = e2e_blocking_ms = None
    e2e_api = None
    if full:
        # e2e: reference-facing calls with HOST buffers (reset -> operator() -> read_result), every step
        e2e_ms = []
        for i in range(nsteps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            eng.reset(hq[i])
            if world == 1:
                eng.run_timed(K)
                v_e, i_e, _ = eng.read_result()
            else:
                sharded.step(stream)
                torch.cuda.synchronize()
                v_e, i_e, _ = eng.read_result()
            dt = (time.perf_counter() - t0) * 1e3
            if i >= warmup:
                e2e_ms.append(dt)
        e2e_t = torch.tensor([sum(e2e_ms) / len(e2e_ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_blocking_ms = float(e2e_t.item())
        # the device-resident and the host-buffer paths must agree on the last query
        assert np.array_equal(i_e, idx_last) and np.array_equal(v_e, val_last), "e2e and resident results differ"
        # e2e, throughput form of the same verbs (tks_submit_host / tks_fetch): host query in, host result out for EVERY
        # step, up to four steps in flight; every step's H2D copy and result read-back happen inside the timed region
        e2e_ms_step = e2e_blocking_ms
        e2e_api = "SpMV.reset(host vec) -> operator() -> read_result(host)"
        if pipelined:
            depth = 3

            def host_loop(lo, hi):
                tickets, last = [], None
                for i in range(lo, hi):
                    tickets.append(sharded.submit_host(hq[i]))
                    if len(tickets) > depth:
                        last = sharded.fetch(tickets[len(tickets) - 1 - depth])
                for t in tickets[max(0, len(tickets) - depth):]:
                    last = sharded.fetch(t)
                return last

            host_loop(0, warmup)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            aligned_start(torch, dist, world)
            t0 = time.perf_counter()
            v_p, i_p, _ = host_loop(warmup, nsteps)
            dt = (time.perf_counter() - t0) * 1e3 / steps
            e2e_t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
            e2e_ms_step = float(e2e_t.item())
            assert np.array_equal(i_p, idx_last) and np.array_equal(v_p, val_last), "pipelined e2e and resident results differ"
            e2e_api = ("SpMV.submit_host(host vec) -> ticket ... SpMV.fetch(ticket) -> host result, for every step; up to 4 steps "
                       "in flight (tks_submit_host / tks_fetch); wall clock from the first submit to the last fetch")

    # roofline of the dominant kernel (csr_topk_main_kernel), timed alone with CUDA events on its stream
    main_ms = measure_main_kernel(tks, eng, hq, warmup, steps if full else min(steps, 5), K)
    achieved = alg_bytes_local / (main_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "csr_topk_main_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs, "traffic": load_traffic(wl_key), "traffic_source": traffic_source(wl_key), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes_local, "main_kernel_ms": main_ms,
            "step_frac": alg_bytes_local / (ms_step * 1e-3) / 1e9 / peak_gbs,
            # the device layout is smaller than the CSR the algorithmic figure counts (16-bit column offsets + a
            # row-start bitmap instead of 32-bit indices + row_ptr): bytes the kernel actually streams, and their rate
            "streamed_bytes_per_launch": int(stats.device_bytes),
            "streamed_gbs": int(stats.device_bytes) / (main_ms * 1e-3) / 1e9,
            "streamed_frac": int(stats.device_bytes) / (main_ms * 1e-3) / 1e9 / peak_gbs,
            "streamed_step_frac": int(stats.device_bytes) / (ms_step * 1e-3) / 1e9 / peak_gbs}

    cpu = None
    if full and rank == 0 and world == 1 and not args.no_cpu and not half:
        cpu = cpu_baseline_leg(tks, eng, hq[warmup:], idx_last, val_last, args)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        exchange_txt = ("the select kernel itself (stores into every rank's IPC window over NVLink, wait, merge: no extra launch)"
                        if sharded.exchange_mode == "peer" else "NCCL all-gather + merge kernel")
        rec = {"metric": "topk_spmv_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": world, "steps": steps,
               "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "weak" if weak else "strong", "vs_baseline": None,
               "dtype": ("bf16 values x bf16 query, f32 products and sums" if wl.get("bf16") else
                         "f16 values x f16 query, f32 products and sums" if half else "f32"), "data": "synthetic",
               "config": {"workload": wl["name"] + (f", weak-scaled: {world} shards of {wl['rows']} rows" if weak and world > 1 else ""),
                          "rows": rows_total, "cols": cols, "nnz": nnz_total, "k": K,
                          "sharding": (f"rows/{world}, K candidates exchanged by " + exchange_txt + " on every rank") if world > 1 else "none",
                          "pipeline": ("tks_submit: consecutive queries overlap (sample of query i+1 beside the main kernel of query i, "
                                       "main kernels chained by programmatic dependent launch, select on its own stream); "
                                       "every query's three kernels complete inside the timed region") if pipelined else "none (stream order)",
                          "l2": "inputs larger than L2 (matrix %.2f GB per GPU vs 126 MB), no flush" % (nnz_local * (6 if half else 8) / 1e9),
                          "work_unit_nnz": int(stats.work_unit_nnz), "work_units": int(stats.work_units),
                          "generator_s": round(gen_s, 2)},
               "per_step": per_step, "parity_n": parity["ok"], "parity": parity,
               "roofline": roof,
               # ours per step: sample + main + select (which also exchanges and merges in peer mode); the NCCL path adds a merge launch
               "gpu_launches": steps * (3 if (world == 1 or sharded.exchange_mode == "peer") else 4),
               "candidates_last_step": int(stats.last_candidates),
               "hbm_gbs_effective": alg_bytes_local * world / (ms_step * 1e-3) / 1e9,
               "clocks": clocks}
        if full:
            rec["cpu_baseline"] = cpu
            rec["e2e"] = {"value": nnz_total / (e2e_ms_step * 1e-3), "unit": "nnz/s", "ms_per_step": e2e_ms_step,
                          "h2d_bytes_per_step": cols * 4, "d2h_bytes_per_step": K * 8 + 4, "api": e2e_api,
                          # the reference hosts' strictly sequential loop (host_spmv_topk_csr_gpu.cu:399-423), one query at a time
                          "blocking": {"value": nnz_total / (e2e_blocking_ms * 1e-3), "ms_per_step": e2e_blocking_ms,
                                       "api": "SpMV.reset(host vec) -> operator() -> read_result(host), one query at a time"}}
    eng.close()
    torch.cuda.empty_cache()
    return rec


def parity_check(tks, torch, dist, eng, sharded, query, query_raw, idx_last, val_last, world, rank, r0, r1, stream, whole_shard):
    """Is the engine's GLOBAL top-k for `query` the reference's?  Every rank runs the reference's CPU gold
    (spmv_coo_gold_top_k, all host threads) over its own shard -- the whole shard when it is small enough to copy
    back (cfg2: 10M rows), else a bounded sample of it plus an exact re-computation of every returned row -- the
    per-rank lists are gathered and merged on the host, and the engine's result must match (north-star tolerance)."""
    out = {"ok": True, "ranks": world}
    mine = None
    rows_local = r1 - r0
    sample_rows = rows_local if whole_shard else min(rows_local, 500_000)
    ptr, idx, val = eng.download_csr_rows(0, sample_rows)
    _, kind, cores, res = cpu_reference(ptr, idx, val, [query], K, max_seconds=1e9)
    gi, gv = res[0]
    mine = (gi.astype(np.int64) + r0, gv)
    gathered = [mine]
    results = [(np.asarray(idx_last, np.int64), np.asarray(val_last, np.float32))]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        results = [None] * world
        dist.all_gather_object(results, (np.asarray(idx_last, np.int64), np.asarray(val_last, np.float32)))
    out["ranks_hold_identical_results"] = all(np.array_equal(r[0], results[0][0]) and np.array_equal(r[1], results[0][1]) for r in results)
    out["ok"] &= out["ranks_hold_identical_results"]
    out["oracle"] = f"reference spmv_coo_gold_top_k ({kind}) on every rank's own shard, {cores} host threads per rank"
    if whole_shard:
        gi, gv = merge_lists(gathered, K)
        ok, detail = lists_agree(idx_last, val_last, gi, gv)
        out.update(detail)
        out["scope"] = f"whole matrix: {world} shards of {rows_local} rows"
        out["ok"] &= ok
    else:
        # (i) no sampled row that clearly beats the engine's k-th score may be missing from the engine's list
        kth = float(val_last[-1])
        gi, gv = merge_lists(gathered, K)
        have = set(np.asarray(idx_last, np.int64).tolist())
        missed = [(int(i), float(v)) for i, v in zip(gi, gv) if v > kth * (1 + 1e-5) + 1e-7 and int(i) not in have]
        # (ii) every returned row this rank owns: its exact fp32 score re-computed on the host from the resident CSR
        worst = 0.0
        for i, v in zip(np.asarray(idx_last, np.int64), np.asarray(val_last, np.float32)):
            if r0 <= i < r1:
                p, c, w = eng.download_csr_rows(int(i - r0), int(i - r0) + 1)
                acc = np.float32(0)
                for cc, ww in zip(c, w):
                    acc = np.float32(acc + np.float32(ww * query[cc]))
                worst = max(worst, abs(float(acc) - float(v)) / max(abs(float(acc)), 1e-30))
        worst_t = torch.tensor([worst], dtype=torch.float64, device="cuda")
        miss_t = torch.tensor([len(missed)], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(worst_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(miss_t)
        # (iii) the exchange + merge: the un-exchanged local lists of all ranks merged on the host = the engine's list
        eng.reset(np.ascontiguousarray(query_raw))
        eng.run_async(K, stream)
        lv, li, lc = eng.read_result()
        locs = [(np.asarray(li[:lc], np.int64), np.asarray(lv[:lc], np.float32))]
        if world > 1:
            locs = [None] * world
            dist.all_gather_object(locs, (np.asarray(li[:lc], np.int64), np.asarray(lv[:lc], np.float32)))
        mi, mv = merge_lists(locs, K)
        same_merge = bool(np.array_equal(mi, np.asarray(idx_last, np.int64)) and np.array_equal(mv, np.asarray(val_last, np.float32)))
        out.update({"scope": f"first {sample_rows} rows of each of the {world} shards ({rows_local} rows per shard) through the gold, "
                             f"every returned row re-computed exactly, exchange + merge against the host merge of the local lists",
                    "sampled_rows_beating_kth_but_missing": int(miss_t.item()),
                    "max_rel_score_error_of_returned_rows": float(worst_t.item()),
                    "device_merge_equals_host_merge_of_local_lists": same_merge})
        out["ok"] &= int(miss_t.item()) == 0 and float(worst_t.item()) <= 1e-5 and same_merge
    out["ok"] = bool(out["ok"])
    return out


def measure_main_kernel(tks, eng, hq, warmup, steps, k):
    """Average duration of the dominant kernel alone (CUDA events recorded by tks_run on its own stream
    right before and after csr_topk_main_kernel; cfg.profile_kernels)."""
    ms = []
    eng.set_profile_kernels(True)
    for i in range(warmup + steps):
        eng.reset(hq[i % len(hq)])
        km, _ = eng.run_timed(k)
        st = eng.stats()
        v = st.last_main_kernel_ms if st.last_main_kernel_ms > 0 else km
        if i >= warmup:
            ms.append(v)
    eng.set_profile_kernels(False)
    return sum(ms) / len(ms)


def load_traffic(wl_key):
    """dram bytes per launch of the dominant kernel from the committed ncu summary, if any (profiles/traffic.json; the
    entry names the capture file and the git revision it came from -- see `traffic_source` in the line)."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            e = json.loads(p.read_text()).get(wl_key)
            return e.get("bytes") if isinstance(e, dict) else e
        except Exception:
            return None
    return None


def traffic_source(wl_key):
    p = ROOT / "profiles" / "traffic.json"
    try:
        e = json.loads(p.read_text()).get(wl_key)
        return {"capture": "profiles/" + e["capture"], "git": e["git"]} if isinstance(e, dict) else None
    except Exception:
        return None


def cpu_baseline_leg(tks, eng, queries, idx_gpu, val_gpu, args):
    """cpu_baseline over the SAME matrix: the reference's CPU gold on all host threads for a bounded number of queries
    (~20 s), and the float64 stand-in of test_cpu.py's sparse_dot_topn call beside it."""
    ptr, idx, val = eng.download_csr()
    rows = ptr.size - 1
    sample_rows = min(rows, args.ref_rows) if args.ref_rows else rows
    e = int(ptr[sample_rows])
    times, kind, cores, _ = cpu_reference(ptr[:sample_rows + 1], idx[:e], val[:e], queries, K, max_seconds=15.0)
    sec = sum(times) / len(times)
    stand = standin_f64(ptr[:sample_rows + 1], idx[:e], val[:e], queries[:5], K, max_seconds=10.0)
    ssec = sum(stand) / len(stand)
    # top-K recall (the reference's "precision", host_spmv_bscsr.cpp:646-648) of the engine's last result against the
    # reference gold over the WHOLE matrix for that query
    _, _, _, full = cpu_reference(ptr, idx, val, queries[-1:], K, max_seconds=1e9)
    gi, gv = full[0]
    inter = len(set(gi.tolist()) & set(np.asarray(idx_gpu).tolist()))
    # rows outside the intersection must be near-ties of the K-th score (fp32 summation order), never real misses
    kth = float(gv[-1])
    miss = [float(v) for i, v in zip(gi.tolist(), gv.tolist()) if i not in set(np.asarray(idx_gpu).tolist())]
    return {"value": e / sec, "unit": "nnz/s", "cores": cores, "kind": kind,
            "sample": f"{'the whole' if sample_rows == rows else 'first ' + str(sample_rows) + ' rows of the'} benchmark matrix "
                      f"({sample_rows} rows, {e} nnz), {len(times)} queries, reference spmv_coo_gold_top_k over {cores} row blocks in {cores} threads",
            "ms_per_query_on_sample": sec * 1e3,
            "stand_in": {"value": e / ssec, "unit": "nnz/s", "cores": 1, "kind": "port", "ms_per_query": ssec * 1e3,
                         "what": "test_cpu.py:91-105 with sparse_dot_topn (absent, un-pinned) replaced by float64 scipy csr @ vec + "
                                 "argpartition top-k (BASELINE.md 3.2), same rows, %d queries; scipy's product runs on one thread" % len(stand)},
            "recall_vs_reference_gold_full_matrix": {"k": K, "precision": inter / K,
                                                     "max_rel_gap_of_missed_rows_to_kth": max([abs(v - kth) / kth for v in miss], default=0.0),
                                                     "max_abs_score_diff": float(np.max(np.abs(np.sort(gv)[::-1] - np.sort(np.asarray(val_gpu))[::-1])))}}


def ours_batched(args, tks, wl, rows_total, peak_gbs, peak_src, world, rank, local, stream):
    import torch
    import torch.distributed as dist
    line = batched_workload(tks, torch, dist, args, wl, rows_total, peak_gbs, peak_src, world, rank, local, stream,
                            steps=args.steps, warmup=args.warmup, full=True)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and line.get("parity_n") is False:
        raise SystemExit("parity check failed (see parity in the line above)")


def batched_workload(tks, torch, dist, args, wl, rows_total, peak_gbs, peak_src, world, rank, local, stream, steps, warmup, full):
    """cfg5: every step scores one batch of 64 queries against the resident matrix (rows sharded over the
    ranks); `value` counts query x non-zero products per second (SURVEY 8d "amortised figure Q*nnz/s")."""
    cols, B = wl["cols"], wl["batch"]
    shards = tks.sharding.plan_row_shards_even(rows_total, world)
    r0, r1 = shards[rank]
    eng = tks.SpMV(num_cols=cols, k=K, device=local, max_batch=B, profile_kernels=True, batch_fma=bool(args.batch_fma))
    t0 = time.perf_counter()
    eng.generate_synthetic(r1 - r0, cols, wl["deg"], wl["dist"], seed=SEED, row_offset=r0)
    gen_s = time.perf_counter() - t0
    nnz_local = int(eng.stats().nnz)
    nnz_t = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz_total = int(nnz_t.item())
    nsteps = warmup + steps
    nsets = min(nsteps, 4)                       # distinct query batches, cycled
    hq = [make_queries(cols, B, seed0=1 + 1000 * i) for i in range(nsets)]
    dq = [torch.from_numpy(q).cuda() for q in hq]
    KMAX = 1024
    sharded = tks.ShardedSpMV(eng, K, batch=B)

    def step(i):
        eng.reset_device(dq[i % nsets].data_ptr(), B, stream)
        sharded.step(stream)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    aligned_start(torch, dist, world)
    sampler.mark("timed")
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    sampler.mark("after")
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    last_set = (nsteps - 1) % nsets
    res_last = [eng.read_result(q) for q in (0, B - 1)]
    parity = batched_parity(tks, torch, dist, eng, hq[last_set], res_last, (0, B - 1), world, rank, r0, r1)

    # e2e: host queries in, host results out, every step
    e2e_ms, main_ms = [], []
    e2e_ms_step = None
    if full:
        for i in range(nsteps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            eng.reset(hq[i % nsets])
            if world == 1:
                eng.run_timed(K)
                main_ms.append(eng.stats().last_main_kernel_ms)
            else:
                sharded.step(stream)
                torch.cuda.synchronize()
            out = [eng.read_result(q) for q in range(B)]
            dt = (time.perf_counter() - t0) * 1e3
            if i >= warmup:
                e2e_ms.append(dt)
        e2e_t = torch.tensor([sum(e2e_ms) / len(e2e_ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_ms_step = float(e2e_t.item())
        for (v0, i0, _), q in zip(res_last, (0, B - 1)):
            assert np.array_equal(out[q][1], i0) and np.array_equal(out[q][0], v0), "e2e and resident results differ"
    if world > 1 or not full:
        # dominant kernel alone: profiled single-rank runs (tks_run brackets it with events)
        main_ms = []
        eng.reset(hq[0])
        for _ in range(3):
            eng.run_timed(K)
            main_ms.append(eng.stats().last_main_kernel_ms)
    sel = main_ms[warmup:] if (world == 1 and full) else main_ms
    main = sum(sel) / len(sel)
    clocks = sampler.stop() if rank == 0 else None
    st = eng.stats()
    alg = int(st.algorithmic_bytes)
    achieved = alg / (main * 1e-3) / 1e9
    sm_clk = (clocks or {}).get("sm_mhz") or 1900.0
    lds_peak_gbs = 148 * 128 * sm_clk * 1e6 / 1e9            # 128 B/clk/SM shared-memory bandwidth
    lds_bytes = nnz_local * B * 4                             # one table word per (query, non-zero)
    # FP32 issue: one multiply and one add per (query, non-zero) (two separate instructions keep the gold's rounding), 32
    # lanes per warp instruction, one warp instruction per scheduler and clock, 4 schedulers per SM
    fp32_bound_ms = nnz_local * B * (1 if args.batch_fma else 2) / 32 / (148 * 4 * sm_clk * 1e6) * 1e3
    roof = {"bound": "hbm", "kernel": "csr_batched_kernel<MAIN>", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs, "traffic": load_traffic("cfg5"), "traffic_source": traffic_source("cfg5"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg, "main_kernel_ms": main,
            "note": "SURVEY 7-H8: the binding resource is shared-memory bandwidth (one 4-byte table word per query x "
                    "non-zero), not HBM; all three bounds are reported",
            "lds": {"achieved": lds_bytes / (main * 1e-3) / 1e9, "peak": lds_peak_gbs, "unit": "GB/s",
                    "frac": lds_bytes / (main * 1e-3) / 1e9 / lds_peak_gbs,
                    "peak_source": "148 SMs x 128 B/clk x sampled SM clock"},
            "fp32_issue": {"bound_ms": fp32_bound_ms, "frac": fp32_bound_ms / main,
                           "what": "multiply + add per (query, non-zero) as warp instructions over 148 SMs x 4 schedulers at the sampled clock"}}
    cpu = None
    if full and rank == 0 and world == 1 and not args.no_cpu:
        ptr, idx, val = eng.download_csr()
        sample_rows = min(ptr.size - 1, (args.ref_rows or 2_000_000) // 4)
        e = int(ptr[sample_rows])
        times, kind, cores, _ = cpu_reference(ptr[:sample_rows + 1], idx[:e], val[:e], hq[0][:8], K, max_seconds=20.0)
        sec = sum(times) / len(times)
        cpu = {"value": e / sec, "unit": "nnz/s", "cores": cores, "kind": kind,
               "sample": f"first {sample_rows} rows ({e} nnz), {len(times)} queries one after another (the reference has no "
                         f"batched mode), spmv_coo_gold_top_k over {cores} row blocks in {cores} threads"}
    line = None
    if rank == 0:
        line = {"metric": "topk_spmv_nnz_per_s", "value": B * nnz_total / (ms_step * 1e-3), "unit": "nnz/s",
                "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "rows": rows_total, "cols": cols, "nnz": nnz_total, "k": K, "queries": B,
                           "value_counts": "queries x non-zeros per second",
                           "arithmetic": "fma" if args.batch_fma else "separate mul/add (bit-identical to the gold)",
                           "sharding": f"rows/{world}" if world > 1 else "none",
                           "l2": "inputs larger than L2 (matrix %.2f GB per GPU vs 126 MB), no flush" % (nnz_local * 8 / 1e9),
                           "generator_s": round(gen_s, 2)},
                "parity_n": parity["ok"], "parity": parity,
                "roofline": roof,
                "gpu_launches": steps * (5 if world == 1 else 6),
                "batched_fallbacks": int(st.batched_fallbacks), "clocks": clocks}
        if full:
            line["cpu_baseline"] = cpu
            line["e2e"] = {"value": B * nnz_total / (e2e_ms_step * 1e-3), "unit": "nnz/s", "ms_per_step": e2e_ms_step,
                           "h2d_bytes_per_step": B * cols * 4, "d2h_bytes_per_step": B * (KMAX * 8 + 4),
                           "api": "SpMV.reset(host [64, cols]) -> operator() -> read_result(q) for every query"}
    eng.close()
    torch.cuda.empty_cache()
    return line


def batched_parity(tks, torch, dist, eng, queries, results, which, world, rank, r0, r1):
    """Batched mode at N ranks against the reference's CPU gold: for the queries `which` of the last batch, (i) no row of
    a bounded sample of every shard that clearly beats the engine's k-th score may be missing from the engine's global
    list, and (ii) every returned row a rank owns is re-computed on the host with the gold's sequential fp32 arithmetic:
    the batched kernel accumulates in the same order, so those scores must be bit-identical."""
    rows_local = r1 - r0
    sample_rows = min(rows_local, 300_000)
    ptr, idx, val = eng.download_csr_rows(0, sample_rows)
    missed, worst = 0, 0.0
    kind, cores = "?", 0
    for (v_e, i_e, _), q in zip(results, which):
        _, kind, cores, res = cpu_reference(ptr, idx, val, [queries[q]], K, max_seconds=1e9)
        gi, gv = res[0]
        kth = float(v_e[-1])
        have = set(np.asarray(i_e, np.int64).tolist())
        missed += sum(1 for i, v in zip(gi.astype(np.int64) + r0, gv) if v > kth * (1 + 1e-5) + 1e-7 and int(i) not in have)
        for i, v in zip(np.asarray(i_e, np.int64), np.asarray(v_e, np.float32)):
            if r0 <= i < r1:
                _, c, w = eng.download_csr_rows(int(i - r0), int(i - r0) + 1)
                acc = np.float32(0)
                for cc, ww in zip(c, w):
                    acc = np.float32(acc + np.float32(ww * queries[q][cc]))
                worst = max(worst, abs(float(acc) - float(v)))
    worst_t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    miss_t = torch.tensor([missed], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(worst_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(miss_t)
    ok = int(miss_t.item()) == 0 and float(worst_t.item()) == 0.0
    return {"ok": bool(ok), "ranks": world, "queries_checked": list(which),
            "oracle": f"reference spmv_coo_gold_top_k ({kind}) on the first {sample_rows} rows of every rank's shard, {cores} threads per rank",
            "sampled_rows_beating_kth_but_missing": int(miss_t.item()),
            "max_abs_score_error_of_returned_rows_vs_sequential_fp32": float(worst_t.item())}


def warmup_steps(args):
    return args.warmup


def lfr_overflow_packets(x, num_rows, partitions, B, LFR):
    """How many 15-entry packets of the BS-CSR stream hold more than LIMITED_FINISHED_ROWS row segments (counted on the
    host from the row ids, partition rule of host_spmv_bscsr.cpp:136-141): from the first such packet on, the reference
    kernel's row counter -- and every later row index of that partition -- is off."""
    rpp = (int(num_rows) + partitions - 1) // partitions
    bounds = np.searchsorted(x, np.arange(partitions + 1, dtype=np.int64) * rpp, side="left")
    total, packets, first_at = 0, 0, []
    for p in range(partitions):
        r = x[bounds[p]:bounds[p + 1]]
        if r.size == 0:
            first_at.append(None)
            continue
        n = (r.size + B - 1) // B
        pad = np.full(n * B, r[-1], r.dtype)
        pad[:r.size] = r
        m = pad.reshape(n, B)
        seg = 1 + (m[:, 1:] != m[:, :-1]).sum(axis=1)
        over = np.nonzero(seg > LFR)[0]
        total += int(over.size)
        packets += int(n)
        first_at.append(float(over[0]) / n if over.size else None)
    hit = [f for f in first_at if f is not None]
    return {"packets": total, "of": packets, "fraction": total / max(packets, 1), "partitions_hit": len(hit),
            "mean_position_of_first_event_in_its_partition": float(np.mean(hit)) if hit else None}


def fixed_workload(args, tks, wl, rows_total, queries, peak_gbs, peak_src, steps, warmup, full):
    """cfg3: the cfg2 matrix quantised to 20-bit fixed point and packed into BS-CSR packets by the host
    packet builder (the reference does this on the host too), 32 partitions x LFR 4 x local K 8."""
    import torch
    W, P, Kp, LFR = 20, 32, 8, 4
    cols = wl["cols"]
    src = tks.SpMV(num_cols=cols, k=K)
    src.generate_synthetic(rows_total, cols, wl["deg"], wl["dist"], seed=SEED)
    ptr, idx, val = src.download_csr()
    deg = np.diff(ptr.astype(np.int64))
    x = np.repeat(np.arange(rows_total, dtype=np.uint32), deg)
    nnz = int(ptr[-1])
    val32 = tks.capi.fixed32_from_double_np(val.astype(np.float64))
    t0 = time.perf_counter()
    eng = tks.SpMVFixed(x, idx, val32, rows_total, cols, k=K, fixed_width=W, partitions=P, local_k=Kp,
                        limited_finished_rows=LFR, device_pack=True)
    pack_s = time.perf_counter() - t0
    q32 = tks.capi.fixed32_from_double_np(queries.astype(np.float64))   # create_sample_vector<real_type_inout> cast
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    dq = torch.from_numpy(q32.view(np.int32)).cuda()

    pipelined = os.environ.get("TKS_BENCH_PIPELINE", "1") != "0"

    def step(i):
        if pipelined:
            eng.submit(dq[i].data_ptr(), K, stream, query_ready=True)   # transform + sample of step i beside stream / replay of step i-1
        else:
            eng.reset_device(dq[i].data_ptr(), stream)
            eng.run_async(K, stream)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sampler.mark("timed")
    for i in range(steps):
        step(warmup + i)
    e1.record()
    sampler.sample_now("timed")                      # the host runs ahead: the device is still inside the timed steps here
    torch.cuda.synchronize()
    sampler.mark("after")
    ms_step = e0.elapsed_time(e1) / steps
    clocks = sampler.stop()
    v_last, i_last = eng.read_result()

    e2e_ms, main_ms = [], []
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.reset(q32[i])
        eng.run_timed(K)
        v_e, i_e = eng.read_result()
        dt = (time.perf_counter() - t0) * 1e3
        if i >= warmup:
            e2e_ms.append(dt)
    assert np.array_equal(i_e, i_last) and np.array_equal(v_e, v_last), "e2e and resident results differ"
    e2e_blocking_ms = sum(e2e_ms) / len(e2e_ms)
    # e2e, throughput form (tks_submit_host / tks_fetch): host query in, host result out (words + host merge) for every
    # step; the sample of step i+1 overlaps the stream and replay kernels of step i, the merge of step i runs on the host
    # while the device works on step i+1
    def host_loop(lo, hi):
        last, prev = None, None
        for i in range(lo, hi):
            t = eng.submit_host(q32[i], K)
            if prev is not None:
                last = eng.fetch(prev)
            prev = t
        return eng.fetch(prev)
    host_loop(0, warmup)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v_p, i_p = host_loop(warmup, warmup + steps)
    e2e_pipe_ms = (time.perf_counter() - t0) * 1e3 / steps
    assert np.array_equal(i_p, i_last) and np.array_equal(v_p, v_last), "pipelined e2e and resident results differ"
    e2e_ms = [e2e_pipe_ms]
    # roofline leg: the dominant kernel bracketed alone (profile_kernels adds two events and a statistics read-back
    # per run, so it is switched on only here)
    eng.set_profile_kernels(True)
    for i in range(warmup + steps):
        eng.reset(q32[i])
        eng.run_timed(K)
        if i >= warmup:
            main_ms.append(eng.stats().last_main_kernel_ms)
    recall = None
    host_pack_s = None
    if full:
        # top-K recall of the approximate design (20-bit fixed point, 32 partitions x local K=8) against the exact fp32
        # engine on the same matrix and queries, with the reference's metrics (plot_errors.py)
        # ... for the reference's semantics (bit-exact, incl. its row-counter drift) and for the engine's
        # drift-free mode (same kernel, same speed; true row indices)
        t0 = time.perf_counter()
        eng_df = tks.SpMVFixed(x, idx, val32, rows_total, cols, k=K, fixed_width=W, partitions=P, local_k=Kp,
                               limited_finished_rows=LFR, drift_free=True)
        host_pack_s = time.perf_counter() - t0
        recalls, recalls_df, df_ms = [], [], []
        for i in range(warmup, warmup + min(steps, 5)):
            src.reset(queries[i])
            src()
            ev, ei, _ = src.read_result()
            eng.reset(q32[i])
            eng.run_timed(K)
            av, ai = eng.read_result()
            recalls.append(tks.accuracy.report(ei, ev, ai, av.astype(np.float64) / 2.0 ** 31))
            eng_df.reset(q32[i])
            km, _ = eng_df.run_timed(K)
            dv, di = eng_df.read_result()
            recalls_df.append(tks.accuracy.report(ei, ev, di, dv.astype(np.float64) / 2.0 ** 31))
            df_ms.append(km)
        eng_df.close()
        mean_of = lambda rs: {k2: float(np.mean([r[k2] for r in rs if k2 in r])) for k2 in rs[0]} if rs else None
        overflow_gamma = lfr_overflow_packets(x, rows_total, P, tks.capi.bscsr_packet_size(W), LFR)
        # The published accuracy of the 20-bit design (errors_2021_03_07.png: 96.7-98.4 % precision at N = 10^7 for K = 8..100)
        # is the reference's "uniform + GloVe" plot (plot_errors.py:38,253-256: KIND "uniform" has its axis floor at 0.96,
        # KIND "gamma" at 0.80): uniform row degrees (10..30 for 20 nnz per row) never put more than LFR = 4 row ends into
        # a 15-entry packet, so the row counter never drifts.  The same engine on a uniform-20 matrix of the same size:
        uniform = None
        if not args.no_uniform:
            src.generate_synthetic(rows_total, cols, wl["deg"], "uniform", seed=SEED)
            uptr, uidx, uval = src.download_csr()
            ux = np.repeat(np.arange(rows_total, dtype=np.uint32), np.diff(uptr.astype(np.int64)))
            uval32 = tks.capi.fixed32_from_double_np(uval.astype(np.float64))
            eng_u = tks.SpMVFixed(ux, uidx, uval32, rows_total, cols, k=K, fixed_width=W, partitions=P, local_k=Kp,
                                  limited_finished_rows=LFR, device_pack=True)
            ur, ums = [], []
            for i in range(warmup, warmup + min(steps, 5)):
                src.reset(queries[i])
                src()
                ev, ei, _ = src.read_result()
                eng_u.reset(q32[i])
                km, _ = eng_u.run_timed(K)
                av, ai = eng_u.read_result()
                ur.append(tks.accuracy.report(ei, ev, ai, av.astype(np.float64) / 2.0 ** 31))
                ums.append(km)
            uniform = {"matrix": f"synthetic {rows_total} x {cols}, uniform ~{wl['deg']} nnz/row, same engine knobs",
                       "reference_semantics": mean_of(ur), "step_ms": float(np.mean(ums[1:])) if len(ums) > 1 else None,
                       "packets_with_more_than_LFR_row_segments": lfr_overflow_packets(ux, rows_total, P, tks.capi.bscsr_packet_size(W), LFR),
                       "published": "FPGA 20-bit, N = 10^7: 96.7-98.4 % precision for K = 8..100 (uniform + GloVe matrices)"}
            eng_u.close()
        src.close()
        recall = {"reference_semantics": mean_of(recalls), "drift_free_mode": mean_of(recalls_df),
                  "packets_with_more_than_LFR_row_segments": overflow_gamma, "uniform_rows": uniform,
                  "drift_free_step_ms": float(np.mean(df_ms[1:])) if len(df_ms) > 1 else None,
                  "note": "precision / Kendall tau / NDCG of plot_errors.py against the exact fp32 engine on the same matrix and "
                          "queries; gamma-distributed rows put more than LFR row segments into ~6e-5 of the packets, after which "
                          "the reference's row counter (and therefore every later row index of the partition) is off by one per event"}
    else:
        src.close()
    e2e_ms_step = sum(e2e_ms) / len(e2e_ms)
    main = sum(main_ms) / len(main_ms)
    st = eng.stats()
    alg = int(st.algorithmic_bytes)
    achieved = alg / (main * 1e-3) / 1e9
    # issue-rate bound: the decode is integer work; report both bounds
    roof = {"bound": "hbm", "kernel": "bscsr_stream_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs, "traffic": load_traffic("cfg3"), "traffic_source": traffic_source("cfg3"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg, "main_kernel_ms": main,
            "step_frac": alg / (ms_step * 1e-3) / 1e9 / peak_gbs}
    cpu = None
    if full and not args.no_cpu:
        sys.path.insert(0, str(ROOT / "oracle"))
        import oracle
        sample_rows = min(rows_total, (args.ref_rows or 2_000_000) // 4)
        e = int(ptr[sample_rows])
        t0 = time.perf_counter()
        packed = oracle.pack_bscsr(x[:e], idx[:e], val32[:e], sample_rows, P, W)
        tp = time.perf_counter() - t0
        times = []
        for qi in range(min(3, len(q32))):
            t0 = time.perf_counter()
            iw, vw = oracle.bscsr_kernel(packed, q32[warmup + qi], Kp, LFR)
            oracle.read_result(iw, vw, packed["first_row"], packed["B"])
            times.append(time.perf_counter() - t0)
        sec = sum(times) / len(times)
        cpu = {"value": e / sec, "unit": "nnz/s", "cores": 1, "kind": "port",
               "sample": f"first {sample_rows} rows ({e} nnz), {len(times)} queries, oracle's sequential transcription of the "
                         f"HLS kernel (the reference has no CPU build of this path); packing took {tp:.1f} s"}
    line = {"metric": "topk_spmv_nnz_per_s", "value": nnz / (ms_step * 1e-3), "unit": "nnz/s", "n_gpus": 1,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 (20-bit fixed point)", "data": "synthetic",
            "config": {"workload": wl["name"], "rows": rows_total, "cols": cols, "nnz": nnz, "k": K,
                       "fixed_width": W, "partitions": P, "local_k": Kp, "limited_finished_rows": LFR,
                       "packets": int(st.packets), "l2": "inputs larger than L2 (%.2f GB of packets), no flush" % (st.packets * 64 / 1e9),
                       "device_pack_upload_s": round(pack_s, 2), "host_pack_upload_s": round(host_pack_s, 2) if host_pack_s is not None else None,
                           "pack": "BS-CSR packets and chunk tables built on the GPU (tks_upload_coo_fixed, H2D of the COO included); "
                                   "host_pack_upload_s is the host packer + upload of the drift-free engine beside it"},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": nnz / (e2e_ms_step * 1e-3), "unit": "nnz/s", "ms_per_step": e2e_ms_step,
                    "h2d_bytes_per_step": cols * 4, "d2h_bytes_per_step": P * Kp * 128,
                    "api": "SpMVFixed.submit_host(host vec) -> ticket ... fetch(ticket) (host merge of P x K x LFR candidates), two steps in flight",
                    "blocking": {"value": nnz / (e2e_blocking_ms * 1e-3), "ms_per_step": e2e_blocking_ms,
                                 "api": "SpMVFixed.reset(host vec) -> operator() -> read_result, one query at a time"}},
            "gpu_launches": steps * 3, "results_returned": int(i_last.size),
            "logged_candidates_last_step": int(st.logged_candidates), "recall_vs_exact_fp32": recall,
            "clocks": clocks}
    line["parity_n"] = True   # the assertions above would have stopped the run otherwise
    line["parity"] = ("the stream-order, pipelined and host-fed paths return identical lists (asserted above); bit-exactness against the "
                      "oracle at this size: tests/test_gpu_full_size.py")
    eng.close()
    torch.cuda.empty_cache()
    return line



def fixed_sharded_workload(tks, torch, dist, args, world, rank, local, steps, warmup):
    """cfg3 over several GPUs (FPGA mode): the reference's 32 row partitions dealt out over the ranks
    (ShardedSpMVFixed), every step = reset(vec) -> operator() on every rank -> all-gather of the result words (32 KB in
    total) -> the reference's host merge on every rank, measured with the blocking verbs and, as the line's value, with
    submit / fetch (two queries in flight).  Strong scaling of the 10M-row matrix."""
    wl = WORKLOADS["cfg3"]
    W, P, Kp, LFR = 20, 32, 8, 4
    rows, cols = wl["rows"], wl["cols"]
    src = tks.SpMV(num_cols=cols, k=K, device=local)
    src.generate_synthetic(rows, cols, wl["deg"], wl["dist"], seed=SEED)     # every rank builds the same matrix and keeps its part
    ptr, idx, val = src.download_csr()
    src.close()
    x = np.repeat(np.arange(rows, dtype=np.uint32), np.diff(ptr.astype(np.int64)))
    nnz = int(ptr[-1])
    val32 = tks.capi.fixed32_from_double_np(val.astype(np.float64))
    del val
    t0 = time.perf_counter()
    s = tks.ShardedSpMVFixed(x, idx, val32, rows, cols, k=K, fixed_width=W, partitions=P, local_k=Kp, limited_finished_rows=LFR,
                             device=local, device_pack=True)
    pack_s = time.perf_counter() - t0
    queries = make_queries(cols, warmup + steps)
    q32 = tks.capi.fixed32_from_double_np(queries.astype(np.float64))
    for i in range(warmup):
        s.reset(q32[i])
        s()
    torch.cuda.synchronize()
    dist.barrier()
    aligned_start(torch, dist, world)
    t0 = time.perf_counter()
    kernel_ns = []
    for i in range(steps):
        s.reset(q32[warmup + i])
        kernel_ns.append(s())
    dt = (time.perf_counter() - t0) * 1e3 / steps
    v_blk, i_blk = s.read_result()
    # throughput form: submit(vec) / fetch(ticket), two queries in flight -- the query copy, transform and sample of step
    # i+1 overlap the kernels of step i, the all-gather and the device-to-host copy follow on the stream, and the host
    # merge of step i runs while the device works on step i+1
    def pipe_loop(lo, hi):
        prev, last = None, None
        for i in range(lo, hi):
            t = s.submit(q32[i])
            if prev is not None:
                last = s.fetch(prev)
            prev = t
        return s.fetch(prev)
    pipe_loop(0, warmup)
    torch.cuda.synchronize()
    dist.barrier()
    aligned_start(torch, dist, world)
    t0 = time.perf_counter()
    v_last, i_last = pipe_loop(warmup, warmup + steps)
    dt_pipe = (time.perf_counter() - t0) * 1e3 / steps
    assert np.array_equal(i_last, i_blk) and np.array_equal(v_last, v_blk), "pipelined and blocking partition exchange differ"
    t = torch.tensor([dt, float(np.mean(kernel_ns)) * 1e-6, dt_pipe], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_blocking, kernel_ms, ms_step = float(t[0].item()), float(t[1].item()), float(t[2].item())
    # parity: every rank holds the same list, and it is the list ONE device holding all 32 partitions returns (rank 0
    # builds that engine beside its shard) -- which tests/test_gpu_full_size.py holds bit-exact to the oracle
    lists = [None] * world
    dist.all_gather_object(lists, (np.asarray(i_last), np.asarray(v_last)))
    same = all(np.array_equal(l[0], lists[0][0]) and np.array_equal(l[1], lists[0][1]) for l in lists)
    one = None
    if rank == 0:
        full = tks.SpMVFixed(x, idx, val32, rows, cols, k=K, fixed_width=W, partitions=P, local_k=Kp,
                             limited_finished_rows=LFR, device=local, device_pack=True)
        full.reset(q32[warmup + steps - 1])
        full()
        fv, fi = full.read_result()
        full.close()
        one = bool(np.array_equal(fi, i_last) and np.array_equal(fv, v_last))
    flag = [one]
    dist.broadcast_object_list(flag, src=0)
    s.close()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"metric": "topk_spmv_nnz_per_s", "value": nnz / (ms_step * 1e-3), "unit": "nnz/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "dtype": "u32 (20-bit fixed point)",
            "data": "synthetic",
            "config": {"workload": wl["name"] + f", the 32 partitions dealt out over {world} GPUs ({P // world} each)", "rows": rows,
                       "cols": cols, "nnz": nnz, "k": K, "fixed_width": W, "partitions": P, "local_k": Kp, "limited_finished_rows": LFR,
                       "step": "submit(host vec) -> kernels on every rank -> all-gather of the result words -> one device-to-host copy "
                               "-> fetch(): host merge on every rank; two queries in flight (wall clock, max over ranks)",
                       "pack_upload_s": round(pack_s, 2)},
            "blocking": {"ms_per_step": ms_blocking, "value": nnz / (ms_blocking * 1e-3),
                         "api": "reset(host vec) -> operator() -> read_result(), one query at a time"},
            "local_kernels_ms": kernel_ms,
            "parity_n": bool(same and flag[0]),
            "parity": {"ranks_hold_identical_results": bool(same), "equals_one_device_holding_all_partitions": flag[0],
                       "oracle": "tests/test_gpu_full_size.py holds the one-device engine bit-exact to the oracle at this size"},
            "gpu_launches": steps * 3}


def ours_fixed(args, tks, wl, rows_total, queries, peak_gbs, peak_src):
    print(json.dumps(fixed_workload(args, tks, wl, rows_total, queries, peak_gbs, peak_src, args.steps, args.warmup, True)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)      # NITER of test_spmv_topk.py:21
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, *WORKLOADS.keys()])
    ap.add_argument("--rows", type=int, default=0, help="override the workload's total rows (debug)")
    ap.add_argument("--ref-rows", type=int, default=0, help="cap on the rows the CPU legs run over (0 = the whole configuration)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="N = 1: skip the cfg3 / cfg5 sub-records of the default line")
    ap.add_argument("--no-uniform", action="store_true", help="cfg3: skip the uniform-rows accuracy leg")
    ap.add_argument("--no-cfg4", action="store_true", help="N > 1: skip the BASELINE config 4 sub-record")
    ap.add_argument("--batch-fma", action="store_true", help="cfg5: fused multiply-add arithmetic")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
