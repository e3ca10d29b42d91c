# round 2, GPU call ai (8 GPUs, final defaults: 8192-non-zero work units): the driver's multi-GPU lines -- pipelined submits with the exchange in the select kernel at
# 8 and 4 ranks (weak-scaled cfg2 + parity_n + cfg4 and cfg5 sub-records); 8-rank parity tests of the exchange
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -k "exchange" 2>&1 | tail -3
PORT=29990
run() { name=$1; n=$2; shift; shift; ( env "$@" timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/r02ai_bench_$name.json 2> gpurun_out/r02ai_bench_$name.err; PORT=$((PORT+1)); }
run n8 8 A=1
run n2 2 A=1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02ai_bench_n1.json 2> gpurun_out/r02ai_bench_n1.err
python - <<'PY'
import json
for m in ["n1","n2","n8"]:
    try:
        j=json.loads(open(f"gpurun_out/r02ai_bench_{m}.json").read().strip().splitlines()[-1]); ps=j["per_step"] or {}
        print(m, round(j["ms_per_step"],4), j["value"], "main_alone", round(j["roofline"]["main_kernel_ms"],4), "frac", round(j["roofline"]["frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "blocking", round(j["e2e"]["blocking"]["ms_per_step"],4), "per_step", ps.get("mean_ms"), ps.get("std_ms"), ps.get("timeline_us"), j["parity_n"], j["parity"].get("set_difference"))
        for k in ("cfg3","cfg4","cfg5"):
            c=j.get(k)
            if c: print("  ", k, round(c["ms_per_step"],4), c["value"], c.get("per_step") and c["per_step"]["mean_ms"], c["parity_n"], (c.get("roofline") or {}).get("main_kernel_ms"), c.get("local_kernels_ms"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02ai_bench_{m}.err").read()[-1500:])
PY
