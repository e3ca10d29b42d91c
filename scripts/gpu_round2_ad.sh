# round 2, GPU call ad (1 GPU): main CTA size of the 16-bit modes inside the pipeline (2 x 10 / 11 / 12 warps)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --workload cfg2h --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02ad_bench_$name.json 2> gpurun_out/r02ad_bench_$name.err; }
run t320 TKS_PIPE_THREADS_16BIT=320
run t352 TKS_PIPE_THREADS_16BIT=352
run t384 TKS_PIPE_THREADS_16BIT=384
python - <<'PY'
import json
for m in ["t320","t352","t384"]:
    try:
        j=json.loads(open(f"gpurun_out/r02ad_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), (ps.get("timeline_us") or {}), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02ad_bench_{m}.err").read()[-800:])
PY
