# round 1, GPU call am: final state on one GPU -- full suite, smoke, every bench line, launch lists
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_am.log 2>&1
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke_am.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_cfg2_am.json 2> gpurun_out/bench_cfg2_am.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_ref_am.json 2> gpurun_out/bench_ref_am.err
( time timeout 600 python bench.py --workload cfg2h --no-cpu ) > gpurun_out/bench_cfg2h_am.json 2> gpurun_out/bench_cfg2h_am.err
( time timeout 900 python bench.py --workload cfg3 --steps 30 --no-cpu ) > gpurun_out/bench_cfg3_am.json 2> gpurun_out/bench_cfg3_am.err
( time timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/bench_cfg5_am.json 2> gpurun_out/bench_cfg5_am.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_am.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg2_am.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3_am.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3_am.log 2>&1
tail -3 gpurun_out/pytest_gpu_am.log
