# round 1, GPU call b: batched-mode tests, full GPU suite, cfg5 bench at reduced and full size, ncu of the
# BS-CSR stream kernel and the batched kernel
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
( time timeout 600 python -m pytest tests/test_gpu_batched.py -x -q ) > gpurun_out/pytest_batched.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_batched.py ) > gpurun_out/pytest_gpu.log 2>&1
( time timeout 600 python bench.py --workload cfg5 --rows 6250000 --steps 10 --no-cpu ) > gpurun_out/bench_cfg5_6M.json 2> gpurun_out/bench_cfg5_6M.err
( time timeout 600 python bench.py --workload cfg5 --rows 6250000 --steps 10 --no-cpu --batch-fma ) > gpurun_out/bench_cfg5_6M_fma.json 2> gpurun_out/bench_cfg5_6M_fma.err
( time timeout 900 python bench.py --workload cfg5 --steps 10 ) > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
( time timeout 900 python bench.py --workload cfg3 --steps 10 --no-cpu ) > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bscsr_stream -s 3 -c 1 -o gpurun_out/prof_bscsr_stream python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_batched_kernel -s 5 -c 1 -o gpurun_out/prof_batched python bench.py --workload cfg5 --rows 6250000 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --rows 6250000 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3.log 2>&1
ls -la gpurun_out
