# round 2, GPU call w (1 GPU): 12-bit column offsets for the k <= 128 single-query kernels (5.625 B/nnz instead of 6.125)
# and no loads for lanes behind a chunk's end -- float suite, cfg2 / cfg2h A/B against 16-bit offsets
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_group.py tests/test_gpu_batched.py tests/test_gpu_tma.py -x -q 2>&1 | tail -4
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02w_bench_$name.json 2> gpurun_out/r02w_bench_$name.err; }
run cfg2 cfg2 A=1
run cfg2_col16 cfg2 TKS_COL12=0
run cfg2h cfg2h A=1
run cfg2h_col16 cfg2h TKS_COL12=0
run cfg2_nopipe cfg2 TKS_BENCH_PIPELINE=0
python - <<'PY'
import json
for m in ["cfg2","cfg2_col16","cfg2_nopipe","cfg2h","cfg2h_col16"]:
    try:
        j=json.loads(open(f"gpurun_out/r02w_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "alg_frac", round(r["frac"],3), "streamed", round(r["streamed_frac"],3), r["streamed_bytes_per_launch"], "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), (ps.get("timeline_us") or {}).get("main_begin_after_previous_main_end"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02w_bench_{m}.err").read()[-800:])
PY
