# round 2, GPU call ba (1 GPU): sampled work units spread evenly over the stream (unit count / sample count need not be a
# whole number any more) -- float tests incl. full-size parity, smoke, cfg2h and cfg2 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_tma.py tests/test_gpu_group.py -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu --no-sub > gpurun_out/r02ba_bench_$name.json 2> gpurun_out/r02ba_bench_$name.err; }
run cfg2h cfg2h A=1
run cfg2 cfg2 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ba_bench_*.json")):
    m=f.split("r02ba_bench_")[1][:-5]
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), j.get("parity_n"), "candidates", j.get("candidates_last_step"), j["config"]["work_unit_nnz"], j["config"]["work_units"])
    except Exception as e: print(m, "ERR", e, open(f[:-5]+".err").read()[-600:])
PY
