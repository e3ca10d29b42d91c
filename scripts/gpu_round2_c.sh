# round 2, GPU call c (1 GPU): pipelined submits after giving every kernel of the path the same (maximal) shared-memory
# carve-out; %globaltimer timeline of the pipeline; A/B of PDL and CTA size
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_gpu_pipeline.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02c_bench_cfg2.json 2> gpurun_out/r02c_bench_cfg2.err
TKS_BENCH_PIPELINE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02c_bench_cfg2_nopipe.json 2>&1
TKS_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02c_bench_cfg2_pipe_nopdl.json 2>&1
TKS_PIPE_THREADS=576 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02c_bench_cfg2_pipe_576.json 2>&1
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02c_bench_cfg2_100steps.json 2>&1
python - <<'PY'
import json
for m in ["cfg2","cfg2_nopipe","cfg2_pipe_nopdl","cfg2_pipe_576","cfg2_100steps"]:
    try:
        j=json.loads(open(f"gpurun_out/r02c_bench_{m}.json").read().strip().splitlines()[-1]); print(m, j["ms_per_step"], j["value"], j.get("e2e",{}).get("ms_per_step"), j["per_step"], j["parity_n"], j["roofline"]["main_kernel_ms"], j["clocks"]["samples"])
    except Exception as e: print(m, "ERR", e)
PY
cat gpurun_out/r02c_bench_cfg2.err | tail -5
