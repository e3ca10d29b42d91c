# round 2, GPU call h (1 GPU): host-fed pipelined API (tks_submit_host / tks_fetch) -- parity tests, the default bench
# line (e2e through it), the reference arm over the whole 10M-row matrix with the float64 stand-in beside it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_csr.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench_cfg2.json 2> gpurun_out/r02h_bench_cfg2.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02h_bench_cfg2.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j["e2e"], j["per_step"], j["parity_n"], j["roofline"]["main_kernel_ms"], j["clocks"]["samples"])
print(j["cpu_baseline"])
r=json.loads(open("gpurun_out/r02h_bench_ref.json").read().strip().splitlines()[-1])
print(r["ms_per_step"], r["value"], r["cpu_baseline"])
PY
tail -3 gpurun_out/r02h_bench_cfg2.err gpurun_out/r02h_bench_ref.err
