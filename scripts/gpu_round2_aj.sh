# round 2, GPU call aj: BS-CSR work-unit size A/B (TKS_BSCSR_CHUNK packets per unit, TKS_BSCSR_TAIL_DIV divisor of the
# units of the last 10 %): bit-exactness of each variant on the golden + engine tests, then the cfg3 workload timed
mkdir -p gpurun_out
for v in "512 4" "512 1" "768 1" "1024 1" "1024 4" "1024 2"; do
  set -- $v
  echo "== chunk $1 tail_div $2"
  TKS_BSCSR_CHUNK=$1 TKS_BSCSR_TAIL_DIV=$2 timeout 600 python -m pytest tests/test_gpu_bscsr.py tests/test_gpu_golden.py tests/test_gpu_pack.py -x -q 2>&1 | tail -2
  TKS_BSCSR_CHUNK=$1 TKS_BSCSR_TAIL_DIV=$2 timeout 300 python bench.py --workload cfg3 --no-uniform --no-cpu --steps 30 --warmup 5 > gpurun_out/r02aj_cfg3_$1_$2.json 2> gpurun_out/r02aj_cfg3_$1_$2.err
  python - "$1" "$2" <<'PY'
import json,sys
c,t=sys.argv[1:3]
try:
    j=json.loads(open(f"gpurun_out/r02aj_cfg3_{c}_{t}.json").read().strip().splitlines()[-1])
    print("cfg3", c, t, "step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["ms_per_step"],4), "roofline", j["roofline"].get("main_kernel_ms"), round(j["roofline"]["frac"],3), j.get("local_kernels_ms"), j["parity"])
except Exception as e: print("ERR", e, open(f"gpurun_out/r02aj_cfg3_{c}_{t}.err").read()[-1200:])
PY
done
