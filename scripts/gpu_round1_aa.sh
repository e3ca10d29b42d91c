# round 1, GPU call aa: final state on one GPU -- full suite, smoke, every bench line, launch lists, full capture of the main kernel
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_aa.log 2>&1
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke_aa.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_cfg2_aa.json 2> gpurun_out/bench_cfg2_aa.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_ref_aa.json 2> gpurun_out/bench_ref_aa.err
( time timeout 600 python bench.py --workload cfg2h --no-cpu ) > gpurun_out/bench_cfg2h_aa.json 2> gpurun_out/bench_cfg2h_aa.err
( time timeout 900 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_aa.json 2> gpurun_out/bench_cfg3_aa.err
( time timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/bench_cfg5_aa.json 2> gpurun_out/bench_cfg5_aa.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_aa.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg2_aa.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_topk_main -s 3 -c 1 -o gpurun_out/prof_csr_main_aa python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg2_aa.log 2>&1
ls -la gpurun_out | tail -15
