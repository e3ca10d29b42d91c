# round 2, GPU call t (1 GPU): the north-star's staging measured -- k <= 128 main kernel streaming through per-warp rings
# of cp.async.bulk + mbarrier (TKS_TMA=1) against direct 256-bit register loads: parity tests, main kernel alone and step
mkdir -p gpurun_out
TKS_TMA=1 timeout 900 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -4
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > gpurun_out/r02t_bench_$name.json 2> gpurun_out/r02t_bench_$name.err; }
run cfg2h_tma cfg2h TKS_TMA=1
run cfg2h_tma_nopipe cfg2h TKS_TMA=1 TKS_BENCH_PIPELINE=0
run cfg2h_ldg_nopipe cfg2h TKS_BENCH_PIPELINE=0
run cfg2_tma cfg2 TKS_TMA=1
run cfg2_tma_nopipe cfg2 TKS_TMA=1 TKS_BENCH_PIPELINE=0
run cfg2_tma_t512 cfg2 TKS_TMA=1 TKS_TMA_THREADS=512 TKS_BENCH_PIPELINE=0
run cfg2h_tma_t320 cfg2h TKS_TMA=1 TKS_TMA_THREADS=320 TKS_BENCH_PIPELINE=0
python - <<'PY'
import json
for m in ["cfg2h_tma","cfg2h_tma_nopipe","cfg2h_ldg_nopipe","cfg2h_tma_t320","cfg2_tma","cfg2_tma_nopipe","cfg2_tma_t512"]:
    try:
        j=json.loads(open(f"gpurun_out/r02t_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), (ps.get("timeline_us") or {}).get("main_begin_after_previous_main_end"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02t_bench_{m}.err").read()[-800:])
PY
