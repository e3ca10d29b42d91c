# round 2, GPU call j (1 GPU): batched kernel with 8 quad streams per warp (8 queries per lane) -- parity tests, cfg5 on
# one GPU; 16-bit modes after the carve-out policy of the main kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_csr.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -5
run() { name=$1; wl=$2; steps=$3; shift; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu > gpurun_out/r02j_bench_$name.json 2> gpurun_out/r02j_bench_$name.err; }
run cfg5 cfg5 5 A=1
run cfg2h cfg2h 20 A=1
run cfg2b cfg2b 20 A=1
run cfg2 cfg2 20 A=1
python - <<'PY'
import json
for m in ["cfg5","cfg2h","cfg2b","cfg2"]:
    try:
        j=json.loads(open(f"gpurun_out/r02j_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "frac", round(r["frac"],3), "streamed", r.get("streamed_frac"), "lds", r.get("lds"), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("timeline_us"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02j_bench_{m}.err").read()[-800:])
PY
