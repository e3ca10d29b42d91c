# round 1, GPU call j: aligned sample pieces + side-by-side merge, single-reduction replay argmin: parity, launch list, bench
mkdir -p gpurun_out
set -x
( time timeout 1200 python -m pytest tests/test_gpu_bscsr.py tests/test_gpu_golden.py tests/test_gpu_host_exe.py -x -q ) > gpurun_out/pytest_bscsr_j.log 2>&1
( timeout 600 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_j.json 2> gpurun_out/bench_cfg3_j.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3_j.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3_j.log 2>&1
ls -la gpurun_out
