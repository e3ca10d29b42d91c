# round 2, GPU call v (1 GPU): the driver's sequence on one GPU -- smoke, default bench line (now with cfg3 / cfg5
# sub-records), reference arm at N = 1 and (one process) for the N = 4 configuration
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02v_bench_cfg2.json 2> gpurun_out/r02v_bench_cfg2.err ) 2>&1 | grep real
( time timeout 900 python bench.py --impl reference --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02v_bench_ref_n4.json 2> gpurun_out/r02v_bench_ref_n4.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02v_bench_cfg2.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["e2e"]["blocking"]["ms_per_step"], j["parity_n"], j["clocks"]["samples"], j["roofline"]["frac"], j["roofline"]["traffic_source"])
for k in ("cfg3","cfg5"):
    c=j[k]; print(k, c["ms_per_step"], c["value"], c["e2e"]["ms_per_step"] if "e2e" in c else None, c["roofline"]["main_kernel_ms"], c["roofline"]["frac"], c.get("parity_n"))
r=json.loads(open("gpurun_out/r02v_bench_ref_n4.json").read().strip().splitlines()[-1])
print(r["ms_per_step"], r["value"], r["cpu_baseline"]["sample"][:200], r["cpu_baseline"]["same_config"])
PY
tail -3 gpurun_out/r02v_bench_cfg2.err; tail -3 gpurun_out/r02v_bench_ref_n4.err
