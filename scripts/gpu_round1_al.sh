# round 1, GPU call al (4 GPUs): weak-scaled cfg2 at N=4 (peer exchange), cfg4 and cfg5 at N=4
mkdir -p gpurun_out
set -x
run() { name=$1; shift; ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 4 "$@" ) > gpurun_out/bench_al_$name.json 2> gpurun_out/bench_al_$name.err; PORT=$((PORT+1)); }
PORT=29530
run cfg2_n4 --steps 50 --warmup 5
run cfg5_n4 --workload cfg5 --steps 5 --warmup 3
python - <<'PY'
import json
for m in ["cfg2_n4","cfg5_n4"]:
    try:
        j=json.loads(open(f"gpurun_out/bench_al_{m}.json").read().strip().splitlines()[-1]); print(m, j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"])
    except Exception as e: print(m, "ERR", e)
PY
