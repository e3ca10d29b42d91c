# round 2, GPU call az (1 GPU): tks_stats reports the work-unit size and count (bench config, sizing-rule test) -- float
# tests, smoke, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_csr.py tests/test_gpu_pipeline.py tests/test_gpu_host_exe.py -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02az_bench_cfg2.json 2> gpurun_out/r02az_bench_cfg2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02az_bench_cfg2.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["parity_n"], j["roofline"]["frac"], j["roofline"]["main_kernel_ms"], j["roofline"]["traffic"], j["roofline"]["traffic_source"], j["config"]["work_unit_nnz"], j["config"]["work_units"], j["clocks"]["samples"])
for k in ("cfg3","cfg5"):
    c=j[k]; print(k, c["ms_per_step"], c["value"], c.get("parity_n"), c["clocks"]["samples"] if "clocks" in c else None)
PY
tail -2 gpurun_out/r02az_bench_cfg2.err
