# round 2, GPU call al (1 GPU): how much L1 the float main kernel wants -- shared-memory carve-out sweep (TKS_CARVEOUT_PCT;
# 43 % = 100 KB, 57 % = 132 KB (what the co-residency rule computes), 72 % = 164 KB, 86 % = 196 KB) on cfg2 / cfg2h,
# pipelined and stream order; work-unit size of the 16-bit mode (512 non-zeros per warp iteration)
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02al_bench_$name.json 2> gpurun_out/r02al_bench_$name.err; }
for pct in 43 57 72 86; do
  run cfg2_c$pct cfg2 TKS_CARVEOUT_PCT=$pct
  run cfg2_c${pct}_nopipe cfg2 TKS_CARVEOUT_PCT=$pct TKS_BENCH_PIPELINE=0
  run cfg2h_c$pct cfg2h TKS_CARVEOUT_PCT=$pct
  run cfg2h_c${pct}_nopipe cfg2h TKS_CARVEOUT_PCT=$pct TKS_BENCH_PIPELINE=0
done
run cfg2h_u4096 cfg2h TKS_CHUNK_NNZ=4096
run cfg2h_u16384 cfg2h TKS_CHUNK_NNZ=16384
run cfg2h_default cfg2h A=1
run cfg2_default cfg2 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02al_bench_*.json")):
    m=f.split("r02al_bench_")[1][:-5]
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]; tl=ps.get("timeline_us") or {}
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), tl.get("sample"), tl.get("main"), tl.get("main_begin_after_previous_main_end"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f[:-5]+".err").read()[-600:])
PY
