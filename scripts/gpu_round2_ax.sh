# round 2, GPU call ax (8 GPUs): the driver's N = 8 and N = 4 lines with the final code (whole-number work units, pipelined
# partition exchange of the FPGA mode)
mkdir -p gpurun_out
PORT=29960
run() { name=$1; n=$2; ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/r02ax_bench_$name.json 2> gpurun_out/r02ax_bench_$name.err; PORT=$((PORT+1)); }
run n8 8
run n4 4
python - <<'PY'
import json
for m in ["n8","n4"]:
    try:
        j=json.loads(open(f"gpurun_out/r02ax_bench_{m}.json").read().strip().splitlines()[-1]); ps=j["per_step"] or {}
        print(m, round(j["ms_per_step"],4), j["value"], "main_alone", round(j["roofline"]["main_kernel_ms"],4), "frac", round(j["roofline"]["frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "per_step", ps.get("mean_ms"), ps.get("std_ms"), j["parity_n"])
        for k in ("cfg3","cfg4","cfg5"):
            c=j.get(k)
            if c: print("  ", k, round(c["ms_per_step"],4), c["value"], c["parity_n"], (c.get("roofline") or {}).get("main_kernel_ms"), c.get("local_kernels_ms"), (c.get("blocking") or {}).get("ms_per_step"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02ax_bench_{m}.err").read()[-1500:])
PY
