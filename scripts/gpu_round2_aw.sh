# round 2, GPU call aw (2 GPUs): final code -- the multi-rank tests at world = 2 (fail-fast collection), then the driver's
# N = 2 line (weak-scaled cfg2 + cfg4 / cfg5 / cfg3 sub-records, cfg3 through the pipelined partition exchange)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_pipeline.py tests/test_gpu_group.py -x -q 2>&1 | tail -3
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29971 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02aw_bench_n2.json 2> gpurun_out/r02aw_bench_n2.err
python - <<'PY'
import json
try:
    j=json.loads(open("gpurun_out/r02aw_bench_n2.json").read().strip().splitlines()[-1]); ps=j["per_step"] or {}
    print("n2", round(j["ms_per_step"],4), j["value"], "main_alone", round(j["roofline"]["main_kernel_ms"],4), "frac", round(j["roofline"]["frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "per_step", ps.get("mean_ms"), ps.get("std_ms"), j["parity_n"])
    for k in ("cfg3","cfg4","cfg5"):
        c=j.get(k)
        if c: print("  ", k, round(c["ms_per_step"],4), c["value"], c["parity_n"], (c.get("roofline") or {}).get("main_kernel_ms"), c.get("local_kernels_ms"), (c.get("blocking") or {}).get("ms_per_step"))
except Exception as e: print("ERR", e, open("gpurun_out/r02aw_bench_n2.err").read()[-1500:])
PY
