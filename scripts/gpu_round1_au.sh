# round 1, GPU call au: last call of the round -- full suite on the final code, one cfg4-sized shard (25M rows x 40)
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_au.log 2>&1
tail -3 gpurun_out/pytest_gpu_au.log
( timeout 300 python bench.py --workload cfg4 --rows 25000000 --steps 20 --warmup 3 --no-cpu ) > gpurun_out/bench_cfg4_shard25M_au.json 2> gpurun_out/bench_cfg4_shard25M_au.err
python -c "
import json;j=json.loads(open('gpurun_out/bench_cfg4_shard25M_au.json').read().strip().splitlines()[-1]);print('cfg4 25M-row shard',j['ms_per_step'],j['roofline']['main_kernel_ms'],j['e2e']['ms_per_step'],j['candidates_last_step'])"
