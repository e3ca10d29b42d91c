# round 1, GPU call aq (8 GPUs): final multi-GPU numbers with the exchange fused into the select kernel
mkdir -p gpurun_out
set -x
run() { name=$1; shift; ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 "$@" ) > gpurun_out/bench_aq_$name.json 2> gpurun_out/bench_aq_$name.err; PORT=$((PORT+1)); }
PORT=29540
run cfg2_n8_peer --steps 50 --warmup 5
run cfg4_n8 --workload cfg4 --steps 20 --warmup 3
run cfg5_n8 --workload cfg5 --steps 5 --warmup 3
python - <<'PY'
import json
for m in ["cfg2_n8_peer","cfg4_n8","cfg5_n8"]:
    try:
        j=json.loads(open(f"gpurun_out/bench_aq_{m}.json").read().strip().splitlines()[-1]); print(m, j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["gpu_launches"])
    except Exception as e: print(m, "ERR", e)
PY
