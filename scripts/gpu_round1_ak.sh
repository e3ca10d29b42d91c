# round 1, GPU call ak: programmatic dependent launch in the BS-CSR pipeline (sample -> stream -> replay), A/B
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests/test_gpu_bscsr.py tests/test_gpu_pack.py tests/test_gpu_golden.py tests/test_gpu_host_exe.py tests/test_gpu_multirank.py -x -q ) > gpurun_out/pytest_gpu_ak.log 2>&1
tail -4 gpurun_out/pytest_gpu_ak.log
for pdl in 1 0 1 0; do
  ( TKS_PDL=$pdl timeout 900 python bench.py --workload cfg3 --steps 30 --no-cpu ) > gpurun_out/bench_cfg3_ak_pdl$pdl.json 2> gpurun_out/bench_cfg3_ak_pdl$pdl.err
  python -c "
import json;j=json.loads(open('gpurun_out/bench_cfg3_ak_pdl$pdl.json').read().strip().splitlines()[-1]);print('cfg3 pdl=$pdl',j['ms_per_step'],j['roofline']['main_kernel_ms'],j['e2e']['ms_per_step'], j['recall_vs_exact_fp32']['drift_free_mode']['precision@100'])"
done
