# round 2, GPU call i (1 GPU): 16-bit value modes with 16 non-zeros per lane (was 8) -- parity tests, cfg2h / cfg2b lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_batched.py -x -q 2>&1 | tail -5
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > gpurun_out/r02i_bench_$name.json 2> gpurun_out/r02i_bench_$name.err; }
run cfg2h cfg2h A=1
run cfg2b cfg2b A=1
run cfg2h_t320 cfg2h TKS_MAIN_THREADS_16BIT=320
run cfg2h_nopipe cfg2h TKS_BENCH_PIPELINE=0
run cfg2 cfg2 A=1
python - <<'PY'
import json
for m in ["cfg2h","cfg2b","cfg2h_t320","cfg2h_nopipe","cfg2"]:
    try:
        j=json.loads(open(f"gpurun_out/r02i_bench_{m}.json").read().strip().splitlines()[-1]); ps=j["per_step"] or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "frac", round(r["frac"],3), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("timeline_us"), j["parity_n"])
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02i_bench_{m}.err").read()[-800:])
PY
