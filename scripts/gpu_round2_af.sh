# round 2, GPU call af (1 GPU): work-unit size of the float main kernel (every chunk start costs three dependent round
# trips: scheduler atomic, chunk table, first loads): 2048 / 4096 / 8192 / 16384 non-zeros, pipelined and stream order
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02af_bench_$name.json 2> gpurun_out/r02af_bench_$name.err; }
run c2048 TKS_CHUNK_NNZ=2048
run c4096 TKS_CHUNK_NNZ=4096
run c8192 TKS_CHUNK_NNZ=8192
run c16384 TKS_CHUNK_NNZ=16384
run c8192_nopipe TKS_CHUNK_NNZ=8192 TKS_BENCH_PIPELINE=0
run c4096_nopipe TKS_CHUNK_NNZ=4096 TKS_BENCH_PIPELINE=0
python - <<'PY'
import json
for m in ["c2048","c4096","c8192","c16384","c4096_nopipe","c8192_nopipe"]:
    try:
        j=json.loads(open(f"gpurun_out/r02af_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "e2e", round(j["e2e"]["ms_per_step"],4), "blocking", round(j["e2e"]["blocking"]["ms_per_step"],4), ps.get("mean_ms"), j.get("parity_n"), j["candidates_last_step"])
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02af_bench_{m}.err").read()[-800:])
PY
