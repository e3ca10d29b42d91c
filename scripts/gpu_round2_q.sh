# round 2, GPU call q (1 GPU): BS-CSR mode pipelined submits (sample of query i+1 beside stream / replay of query i,
# result words written to pinned host memory) -- bit-exact tests, cfg3 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bscsr.py tests/test_gpu_golden.py -x -q 2>&1 | tail -4
timeout 900 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-uniform > gpurun_out/r02q_bench_cfg3.json 2> gpurun_out/r02q_bench_cfg3.err
TKS_BENCH_PIPELINE=0 timeout 900 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-uniform --no-cpu > gpurun_out/r02q_bench_cfg3_nopipe.json 2>&1
python - <<'PY'
import json
for m in ["cfg3","cfg3_nopipe"]:
    j=json.loads(open(f"gpurun_out/r02q_bench_{m}.json").read().strip().splitlines()[-1])
    print(m, j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["e2e"].get("blocking",{}).get("ms_per_step"), j["roofline"]["main_kernel_ms"], j["roofline"]["frac"], j["roofline"]["step_frac"])
PY
tail -3 gpurun_out/r02q_bench_cfg3.err
