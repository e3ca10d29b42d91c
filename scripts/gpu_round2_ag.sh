# round 2, GPU call ag (1 GPU): 8192-non-zero work units with a tail of 2048 (default for large matrices) -- float suite,
# cfg2 / cfg2h pipelined and stream order, A/B without the tail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_batched.py tests/test_gpu_group.py -x -q 2>&1 | tail -3
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02ag_bench_$name.json 2> gpurun_out/r02ag_bench_$name.err; }
run cfg2 cfg2 A=1
run cfg2_notail cfg2 TKS_CHUNK_TAIL=0
run cfg2_nopipe cfg2 TKS_BENCH_PIPELINE=0
run cfg2_nopipe_notail cfg2 TKS_BENCH_PIPELINE=0 TKS_CHUNK_TAIL=0
run cfg2h cfg2h A=1
run cfg2_16k cfg2 TKS_CHUNK_NNZ=16384
python - <<'PY'
import json
for m in ["cfg2","cfg2_notail","cfg2_16k","cfg2_nopipe","cfg2_nopipe_notail","cfg2h"]:
    try:
        j=json.loads(open(f"gpurun_out/r02ag_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "blocking", round(j["e2e"]["blocking"]["ms_per_step"],4), ps.get("mean_ms"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02ag_bench_{m}.err").read()[-800:])
PY
