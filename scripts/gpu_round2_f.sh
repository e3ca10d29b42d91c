# round 2, GPU call f (2 GPUs): pipelined submits with the exchange fused into the select kernel -- multirank parity
# tests, the driver's N=2 bench (weak-scaled cfg2 + parity_n + cfg4 sub-record), A/B against stream order
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_multirank.py -x -q 2>&1 | tail -5
PORT=29600
run() { name=$1; shift; ( env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02f_bench_$name.json 2> gpurun_out/r02f_bench_$name.err; PORT=$((PORT+1)); }
run n2 A=1
run n2_nopipe TKS_BENCH_PIPELINE=0
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
python - <<'PY'
import json
for m in ["n1","n2","n2_nopipe"]:
    try:
        j=json.loads(open(f"gpurun_out/r02f_bench_{m}.json").read().strip().splitlines()[-1]); ps=j["per_step"] or {}
        print(m, round(j["ms_per_step"],4), j["value"], "main_alone", round(j["roofline"]["main_kernel_ms"],4), "frac", round(j["roofline"]["frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "std", ps.get("std_ms"), ps.get("timeline_us"), j["parity_n"], j["parity"])
        c=j.get("cfg4")
        if c: print("   cfg4", round(c["ms_per_step"],4), c["value"], c["per_step"], c["parity_n"], c["parity"], c["roofline"]["main_kernel_ms"], c["config"]["generator_s"])
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02f_bench_{m}.err").read()[-1500:])
PY
