# round 2, GPU call d (1 GPU): why does the sample of query i+1 not run beside the main kernel of query i?
# lean select (512 threads, 16 KB) + 9 KB sample CTAs; sweep of main CTA size, sample CTA size and L1/shared split
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py -x -q 2>&1 | tail -5
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02d_bench_$name.json 2> gpurun_out/r02d_bench_$name.err; }
run default A=1
run t480 TKS_PIPE_THREADS=480
run t448 TKS_PIPE_THREADS=448
run t448_s64 TKS_PIPE_THREADS=448 TKS_PIPE_SAMPLE_THREADS=64
run c57 TKS_CARVEOUT_PCT=57
run c57_t448 TKS_CARVEOUT_PCT=57 TKS_PIPE_THREADS=448
run c72_t448 TKS_CARVEOUT_PCT=72 TKS_PIPE_THREADS=448
run c44 TKS_CARVEOUT_PCT=44
run nopdl TKS_PDL=0
run nopipe TKS_BENCH_PIPELINE=0
python - <<'PY'
import json
for m in ["default","t480","t448","t448_s64","c57","c57_t448","c72_t448","c44","nopdl","nopipe"]:
    try:
        j=json.loads(open(f"gpurun_out/r02d_bench_{m}.json").read().strip().splitlines()[-1]); ps=j["per_step"] or {}
        print(m, round(j["ms_per_step"],4), "main_alone", round(j["roofline"]["main_kernel_ms"],4), "e2e", round(j["e2e"]["ms_per_step"],4), "std", ps.get("std_ms"), ps.get("timeline_us"), j["parity_n"])
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02d_bench_{m}.err").read()[-500:])
PY
