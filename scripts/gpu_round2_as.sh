# round 2, GPU call as (1 GPU): whole number of work units per stream for batched handles too -- batched tests, cfg5 at the
# full 50M rows and at the 6.25M rows one of 8 ranks holds, each against the fixed 4096-non-zero units
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_csr.py -x -q 2>&1 | tail -3
run() { name=$1; shift; rows=$1; shift; env "$@" timeout 900 python bench.py --workload cfg5 --rows $rows --steps 5 --warmup 3 --no-cpu > gpurun_out/r02as_bench_$name.json 2> gpurun_out/r02as_bench_$name.err; }
run cfg5 50000000 A=1
run cfg5_u4096 50000000 TKS_CHUNK_NNZ=4096
run cfg5_rank_of_8 6250000 A=1
run cfg5_rank_of_8_u4096 6250000 TKS_CHUNK_NNZ=4096
run cfg5_rank_of_4 12500000 A=1
run cfg5_rank_of_4_u4096 12500000 TKS_CHUNK_NNZ=4096
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02as_bench_*.json")):
    m=f.split("r02as_bench_")[1][:-5]
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main", r.get("main_kernel_ms"), "frac", round(r["frac"],3), j.get("parity_n"), j.get("parity"))
    except Exception as e: print(m, "ERR", e, open(f[:-5]+".err").read()[-600:])
PY
