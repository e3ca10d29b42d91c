# round 1, GPU call r: histogram-threshold selection (tau derived inside the main kernel, one-pass select + rank sort)
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r.log 2>&1
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_r.json 2> gpurun_out/bench_cfg2_r.err
( time timeout 600 python bench.py --workload cfg2h --no-cpu ) > gpurun_out/bench_cfg2h_r.json 2> gpurun_out/bench_cfg2h_r.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2_r.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg2_r.log 2>&1
ls -la gpurun_out
