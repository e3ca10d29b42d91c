# round 2, GPU call y (2 GPUs): the driver's N = 2 line with the final code (12-bit column offsets, aligned start, cfg3 /
# cfg4 / cfg5 sub-records), and cfg3 alone over 2 GPUs
mkdir -p gpurun_out
PORT=29800
run() { name=$1; shift; ( timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 20 --warmup 5 "$@" ) > gpurun_out/r02y_bench_$name.json 2> gpurun_out/r02y_bench_$name.err; PORT=$((PORT+1)); }
( time run n2 ) 2>&1 | grep real
run n2_again --no-cfg4
python - <<'PY'
import json
for m in ["n2","n2_again"]:
    try:
        j=json.loads(open(f"gpurun_out/r02y_bench_{m}.json").read().strip().splitlines()[-1]); ps=j["per_step"] or {}
        print(m, round(j["ms_per_step"],4), j["value"], "main_alone", round(j["roofline"]["main_kernel_ms"],4), "e2e", round(j["e2e"]["ms_per_step"],4), "per_step", ps.get("mean_ms"), ps.get("std_ms"), j["parity_n"], j["clocks"]["samples"])
        for k in ("cfg3","cfg4","cfg5"):
            c=j.get(k)
            if c: print("  ", k, round(c["ms_per_step"],4), c["value"], c["parity_n"], c.get("local_kernels_ms"), (c.get("roofline") or {}).get("main_kernel_ms"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02y_bench_{m}.err").read()[-1500:])
PY
