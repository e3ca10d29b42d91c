# round 2, GPU call ak (1 GPU): 32 interleaved query copies in shared memory (conflict-free gathers, one CTA per SM,
# TKS_XCOPIES=1) against the default main kernel -- parity tests of the variant, then cfg2 / cfg2h pipelined and stream order
mkdir -p gpurun_out
TKS_XCOPIES=1 timeout 900 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02ak_bench_$name.json 2> gpurun_out/r02ak_bench_$name.err; }
run cfg2 cfg2 A=1
run cfg2_x1024 cfg2 TKS_XCOPIES=1
run cfg2_x768 cfg2 TKS_XCOPIES=1 TKS_XCOPIES_THREADS=768
run cfg2_x896 cfg2 TKS_XCOPIES=1 TKS_XCOPIES_THREADS=896
run cfg2_nopipe cfg2 TKS_BENCH_PIPELINE=0
run cfg2_x1024_nopipe cfg2 TKS_XCOPIES=1 TKS_BENCH_PIPELINE=0
run cfg2h cfg2h A=1
run cfg2h_x768 cfg2h TKS_XCOPIES=1
run cfg2h_x640 cfg2h TKS_XCOPIES=1 TKS_XCOPIES_THREADS=640
run cfg2h_x768_nopipe cfg2h TKS_XCOPIES=1 TKS_BENCH_PIPELINE=0
run cfg2h_nopipe cfg2h TKS_BENCH_PIPELINE=0
python - <<'PY'
import json
for m in ["cfg2","cfg2_x1024","cfg2_x896","cfg2_x768","cfg2_nopipe","cfg2_x1024_nopipe","cfg2h","cfg2h_x768","cfg2h_x640","cfg2h_nopipe","cfg2h_x768_nopipe"]:
    try:
        j=json.loads(open(f"gpurun_out/r02ak_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "blocking", round(j["e2e"]["blocking"]["ms_per_step"],4), ps.get("mean_ms"), ps.get("timeline_us"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02ak_bench_{m}.err").read()[-800:])
PY
