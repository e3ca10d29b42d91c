# round 1, GPU call p: re-baseline after the container was re-created -- full GPU suite, smoke, all bench lines,
# launch lists and full captures of the current dominant kernels (cfg2 main, cfg3 stream)
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_p.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_p.log 2>&1
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke_p.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_cfg2_p.json 2> gpurun_out/bench_cfg2_p.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_ref_p.json 2> gpurun_out/bench_ref_p.err
( time timeout 900 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_p.json 2> gpurun_out/bench_cfg3_p.err
( time timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/bench_cfg5_p.json 2> gpurun_out/bench_cfg5_p.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_p.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg2_p.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3_p.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3_p.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_topk_main -s 3 -c 1 -o gpurun_out/prof_csr_main_p python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg2_p.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bscsr_stream -s 3 -c 1 -o gpurun_out/prof_bscsr_stream_p python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg3_p.log 2>&1
ls -la gpurun_out
