mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
( time timeout 900 python bench.py --workload cfg3 --steps 10 ) > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_topk_main -s 3 -c 2 -o gpurun_out/prof_csr_main python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg2.log 2>&1
ls -la gpurun_out
