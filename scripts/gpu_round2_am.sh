# round 2, GPU call am (1 GPU): 16384-non-zero work units as the default of the 16-bit value modes -- float suite + the
# full-size parity (now with a cfg2h leg), cfg2h / cfg2b with 16384 (default) / 24576 / 32768, cfg2 untouched
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_tma.py -x -q 2>&1 | tail -3
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02am_bench_$name.json 2> gpurun_out/r02am_bench_$name.err; }
run cfg2h cfg2h A=1
run cfg2b cfg2b A=1
run cfg2h_u24576 cfg2h TKS_CHUNK_NNZ=24576
run cfg2h_u32768 cfg2h TKS_CHUNK_NNZ=32768
run cfg2h_u12288 cfg2h TKS_CHUNK_NNZ=12288
run cfg2 cfg2 A=1
run cfg2_u12288 cfg2 TKS_CHUNK_NNZ=12288
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02am_bench_*.json")):
    m=f.split("r02am_bench_")[1][:-5]
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]; tl=ps.get("timeline_us") or {}
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), tl.get("sample"), tl.get("main"), tl.get("main_begin_after_previous_main_end"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f[:-5]+".err").read()[-600:])
PY
