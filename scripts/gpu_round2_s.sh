# round 2, GPU call s (1 GPU): 16-bit modes back at 72 registers (2 x 12 warps; 2 x 10 in the pipeline) with 128-thread
# sample CTAs (the sample kernel is capped at 64 registers)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > gpurun_out/r02s_bench_$name.json 2> gpurun_out/r02s_bench_$name.err; }
run cfg2h cfg2h A=1
run cfg2b cfg2b A=1
run cfg2h_s64 cfg2h TKS_PIPE_SAMPLE_THREADS=64
python - <<'PY'
import json
for m in ["cfg2h","cfg2b","cfg2h_s64"]:
    try:
        j=json.loads(open(f"gpurun_out/r02s_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "blocking", round(j["e2e"]["blocking"]["ms_per_step"],4), ps.get("timeline_us"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02s_bench_{m}.err").read()[-800:])
PY
