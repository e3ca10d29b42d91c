# round 2, GPU call r (1 GPU): CUDA graph for the blocking verbs (one graph launch instead of three kernel launches), 16-bit
# main kernel capped at 64 registers (2 x 16 warps) -- whole GPU suite, cfg2 / cfg2h / cfg2b lines, A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > gpurun_out/r02r_bench_$name.json 2> gpurun_out/r02r_bench_$name.err; }
run cfg2 cfg2 A=1
run cfg2_nograph cfg2 TKS_GRAPH=0
run cfg2h cfg2h A=1
run cfg2h_t384 cfg2h TKS_MAIN_THREADS_16BIT=384 TKS_PIPE_THREADS_16BIT=320
run cfg2b cfg2b A=1
python - <<'PY'
import json
for m in ["cfg2","cfg2_nograph","cfg2h","cfg2h_t384","cfg2b"]:
    try:
        j=json.loads(open(f"gpurun_out/r02r_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "blocking", round(j["e2e"]["blocking"]["ms_per_step"],4), (ps.get("timeline_us") or {}).get("main_begin_after_previous_main_end"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02r_bench_{m}.err").read()[-800:])
PY
