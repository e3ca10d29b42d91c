# round 2, GPU call ab (1 GPU): bulk L2 prefetch of a later iteration by lane 0 (cp.async.bulk.prefetch.L2), depth 2..4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_csr.py -x -q -k "cfg1 or k_sweep or chunk" 2>&1 | tail -2
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-sub > gpurun_out/r02ab_bench_$name.json 2> gpurun_out/r02ab_bench_$name.err; }
run base cfg2 A=1
run pf2 cfg2 TKS_L2PF=2
run pf3 cfg2 TKS_L2PF=3
run pf4 cfg2 TKS_L2PF=4
run h_base cfg2h A=1
run h_pf2 cfg2h TKS_L2PF=2
run h_pf3 cfg2h TKS_L2PF=3
python - <<'PY'
import json
for m in ["base","pf2","pf3","pf4","h_base","h_pf2","h_pf3"]:
    try:
        j=json.loads(open(f"gpurun_out/r02ab_bench_{m}.json").read().strip().splitlines()[-1]); ps=j.get("per_step") or {}; r=j["roofline"]
        print(m, round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "streamed", round(r["streamed_frac"],3), "e2e", round(j["e2e"]["ms_per_step"],4), ps.get("mean_ms"), j.get("parity_n"))
    except Exception as e: print(m, "ERR", e, open(f"gpurun_out/r02ab_bench_{m}.err").read()[-800:])
PY
