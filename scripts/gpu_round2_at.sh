# round 2, GPU call at (1 GPU): BS-CSR work units sized to a whole number per resident warp (3552 warps, 3659 packets each:
# 8 x 480, 7 x 544, 6 x 640 packets) without the quarter-size tail, against the default 512 + tail
mkdir -p gpurun_out
for v in "512 4" "480 1" "544 1" "640 1" "736 1"; do
  set -- $v
  TKS_BSCSR_CHUNK=$1 TKS_BSCSR_TAIL_DIV=$2 timeout 300 python bench.py --workload cfg3 --no-uniform --no-cpu --steps 30 --warmup 5 > gpurun_out/r02at_cfg3_$1_$2.json 2> gpurun_out/r02at_cfg3_$1_$2.err
  python - "$1" "$2" <<'PY'
import json,sys
c,t=sys.argv[1:3]
try:
    j=json.loads(open(f"gpurun_out/r02at_cfg3_{c}_{t}.json").read().strip().splitlines()[-1])
    print("cfg3", c, t, "step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["ms_per_step"],4), "stream kernel", round(j["roofline"]["main_kernel_ms"],4), round(j["roofline"]["frac"],3))
except Exception as e: print("ERR", e, open(f"gpurun_out/r02at_cfg3_{c}_{t}.err").read()[-1200:])
PY
done
