# round 2, GPU call ay (1 GPU): one rolled emission site per warp iteration in the pool sink (main kernel 30 320 -> 7 384
# SASS instructions, half 48 664 -> 6 240) -- float suite incl. full-size parity, then cfg2 / cfg2h / cfg2b lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_group.py tests/test_gpu_tma.py tests/test_gpu_batched.py -x -q 2>&1 | tail -3
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu --no-sub > gpurun_out/r02ay_bench_$name.json 2> gpurun_out/r02ay_bench_$name.err; }
run cfg2 cfg2 A=1
run cfg2h cfg2h A=1
run cfg2b cfg2b A=1
run cfg2_r2 cfg2 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ay_bench_*.json")):
    m=f.split("r02ay_bench_")[1][:-5]
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "frac", round(r["frac"],3), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), "blocking", round(j["e2e"]["blocking"]["ms_per_step"],4), ps.get("mean_ms"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f[:-5]+".err").read()[-600:])
PY
