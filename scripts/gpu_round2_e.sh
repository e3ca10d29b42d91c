# round 2, GPU call e (1 GPU): stream priorities sample > select > main (the pending select CTA stood in front of the
# sample CTAs); A/B: equal priorities, 256-thread select, CTA sizes
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02e_bench_$name.json 2> gpurun_out/r02e_bench_$name.err; }
run default A=1
run sameprio TKS_PIPE_SELECT_PRIO_DELTA=0
run sel256 TKS_PIPE_SELECT_THREADS=256
run sel256_sameprio TKS_PIPE_SELECT_THREADS=256 TKS_PIPE_SELECT_PRIO_DELTA=0
run sel256_t448 TKS_PIPE_SELECT_THREADS=256 TKS_PIPE_THREADS=448
run t448 TKS_PIPE_THREADS=448
run t576 TKS_PIPE_THREADS=576
run nopdl TKS_PDL=0
python - <<'PY'
import json
for m in ["default","sameprio","sel256","sel256_sameprio","sel256_t448","t448","t576","nopdl"]:
    try:
        j=json.loads(open(f"gpurun_out/r02e_bench_{m}.json").read().strip().splitlines()[-1]); ps=j["per_step"] or {}
        print(m, round(j["ms_per_step"],4), "main_alone", round(j["roofline"]["main_kernel_ms"],4), "e2e", round(j["e2e"]["ms_per_step"],4), "std", ps.get("std_ms"), ps.get("timeline_us"), j["parity_n"])
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02e_bench_{m}.err").read()[-500:])
PY
