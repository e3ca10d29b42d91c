# round 1, GPU call af: replay kernel with parallel pruning of the candidate logs -- parity tests, cfg3 bench, launch list
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests/test_gpu_bscsr.py tests/test_gpu_pack.py tests/test_gpu_golden.py tests/test_gpu_host_exe.py -x -q ) > gpurun_out/pytest_gpu_af.log 2>&1
tail -5 gpurun_out/pytest_gpu_af.log
( time timeout 900 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_af.json 2> gpurun_out/bench_cfg3_af.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg3_af.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3_af.log 2>&1
python -c "
import json;j=json.loads(open('gpurun_out/bench_cfg3_af.json').read().strip().splitlines()[-1]);print('cfg3',j['ms_per_step'],j['roofline']['main_kernel_ms'],j['e2e']['ms_per_step'], j['recall_vs_exact_fp32']['drift_free_mode']['precision@100'])"
