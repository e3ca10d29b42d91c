# round 2, GPU call p (1 GPU): cfg3 line with the uniform-rows accuracy leg and the LFR-overflow packet counts
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg3 --steps 20 --warmup 5 > gpurun_out/r02p_bench_cfg3.json 2> gpurun_out/r02p_bench_cfg3.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02p_bench_cfg3.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["roofline"]["main_kernel_ms"], j["roofline"]["frac"])
r=j["recall_vs_exact_fp32"]
print({k:v for k,v in r["reference_semantics"].items() if k.startswith("precision")})
print({k:v for k,v in r["drift_free_mode"].items() if k.startswith("precision")})
print(r["packets_with_more_than_LFR_row_segments"])
print(r["uniform_rows"])
PY
tail -3 gpurun_out/r02p_bench_cfg3.err
