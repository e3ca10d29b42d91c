# round 1, GPU call l: 6.125 B/nnz float layout (u16 column offsets + row-start bitmap), replay tiles of 4096 chunks
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_l.log 2>&1
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_l.json 2> gpurun_out/bench_cfg2_l.err
( timeout 600 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_l.json 2> gpurun_out/bench_cfg3_l.err
( timeout 900 python bench.py --workload cfg5 --steps 10 --no-cpu ) > gpurun_out/bench_cfg5_l.json 2> gpurun_out/bench_cfg5_l.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_topk_main -s 3 -c 1 -o gpurun_out/prof_csr_main_l python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg2_l.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3_l.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3_l.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2_l.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg2_l.log 2>&1
ls -la gpurun_out
