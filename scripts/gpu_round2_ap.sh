# round 2, GPU call ap (1 GPU): fp32 work-unit size, same box, three alternating repeats of 8192 / 10240 / 12288 non-zeros
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu --no-sub > gpurun_out/r02ap_bench_$name.json 2> gpurun_out/r02ap_bench_$name.err; }
for rep in 1 2 3; do
  for u in 8192 10240 12288; do run cfg2_u${u}_r$rep cfg2 TKS_CHUNK_NNZ=$u; done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ap_bench_*.json")):
    m=f.split("r02ap_bench_")[1][:-5]
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f[:-5]+".err").read()[-600:])
PY
