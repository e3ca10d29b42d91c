# round 2, GPU call x (1 GPU): does the BS-CSR stream kernel lose speed when all of the SM's memory is carved out as
# shared memory (what a replay / sample CTA beside it would need)?
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-uniform > gpurun_out/r02x_bench_$name.json 2> gpurun_out/r02x_bench_$name.err; }
run default A=1
run carve100 TKS_BSCSR_CARVEOUT=100
run carve86 TKS_BSCSR_CARVEOUT=86
python - <<'PY'
import json
for m in ["default","carve100","carve86"]:
    try:
        j=json.loads(open(f"gpurun_out/r02x_bench_{m}.json").read().strip().splitlines()[-1]); r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "stream_alone", round(r["main_kernel_ms"],4), "frac", round(r["frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02x_bench_{m}.err").read()[-800:])
PY
