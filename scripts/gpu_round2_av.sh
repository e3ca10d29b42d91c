# round 2, GPU call av (1 GPU): the driver's sequence with the final defaults of the round (whole-number work units, pipelined FPGA-mode exchange) -- whole GPU suite, smoke, default bench line
# (cfg3 / cfg5 sub-records), cfg5 alone
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02av_bench_cfg2.json 2> gpurun_out/r02av_bench_cfg2.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02av_bench_cfg2.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["e2e"]["blocking"]["ms_per_step"], j["parity_n"], j["clocks"]["samples"], j["roofline"]["frac"], j["roofline"]["main_kernel_ms"], j["per_step"]["mean_ms"])
print(j["cpu_baseline"]["value"], j["cpu_baseline"]["stand_in"]["value"], j["cpu_baseline"]["recall_vs_reference_gold_full_matrix"])
for k in ("cfg3","cfg5"):
    c=j[k]; print(k, c["ms_per_step"], c["value"], c["e2e"]["ms_per_step"] if "e2e" in c else None, c["roofline"]["main_kernel_ms"], c["roofline"]["frac"], (c["roofline"].get("lds") or {}).get("frac"), c.get("parity_n"))
PY
tail -3 gpurun_out/r02av_bench_cfg2.err
