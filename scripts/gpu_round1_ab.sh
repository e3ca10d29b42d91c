# round 1, GPU call ab: 18-warp CTAs for the k <= 128 main kernel (A/B against 16), results written straight to the
# pinned host block in tks_run, profile events only in the roofline leg
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_ab.log 2>&1
for t in 576 512 576 512; do
  ( TKS_MAIN_THREADS=$t timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_ab_$t.json 2> gpurun_out/bench_cfg2_ab_$t.err
  python -c "
import json;j=json.loads(open('gpurun_out/bench_cfg2_ab_$t.json').read().strip().splitlines()[-1]);print('cfg2 $t',j['ms_per_step'],j['roofline']['main_kernel_ms'],j['e2e']['ms_per_step'])" >> gpurun_out/ab_summary.txt
  ( TKS_MAIN_THREADS=$t timeout 600 python bench.py --no-cpu --workload cfg2h ) > gpurun_out/bench_cfg2h_ab_$t.json 2> gpurun_out/bench_cfg2h_ab_$t.err
  python -c "
import json;j=json.loads(open('gpurun_out/bench_cfg2h_ab_$t.json').read().strip().splitlines()[-1]);print('cfg2h $t',j['ms_per_step'],j['roofline']['main_kernel_ms'],j['e2e']['ms_per_step'])" >> gpurun_out/ab_summary.txt
done
( time timeout 900 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_ab.json 2> gpurun_out/bench_cfg3_ab.err
cat gpurun_out/ab_summary.txt
