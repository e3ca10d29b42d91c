# round 2, GPU call aq (1 GPU): fp32 work-unit size at finer steps (same box as each other; r02ap showed +-0.1 % repeatability)
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu --no-sub > gpurun_out/r02aq_bench_$name.json 2> gpurun_out/r02aq_bench_$name.err; }
for u in 8192 9216 11264 12288 13312 14336 15360 16384 20480 24576; do run cfg2_u$u cfg2 TKS_CHUNK_NNZ=$u; done
for u in 16384 20480 24576 28672; do run cfg2h_u$u cfg2h TKS_CHUNK_NNZ=$u; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02aq_bench_*.json"), key=lambda f:(f.split("_u")[0], int(f.split("_u")[1][:-5]))):
    m=f.split("r02aq_bench_")[1][:-5]
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f[:-5]+".err").read()[-600:])
PY
