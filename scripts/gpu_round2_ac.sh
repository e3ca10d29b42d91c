# round 2, GPU call ac (1 GPU): after compiling the L2-prefetch experiment out -- 16-bit and fp32 lines back where they were?
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02ac_bench_$name.json 2> gpurun_out/r02ac_bench_$name.err; }
run cfg2 cfg2 A=1
run cfg2h cfg2h A=1
run cfg2b cfg2b A=1
python - <<'PY'
import json
for m in ["cfg2","cfg2h","cfg2b"]:
    try:
        j=json.loads(open(f"gpurun_out/r02ac_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), (ps.get("timeline_us") or {}), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02ac_bench_{m}.err").read()[-800:])
PY
