# round 1, GPU call h: host executable test + full suite, cfg3 log statistics, ncu of the small BS-CSR kernels
mkdir -p gpurun_out
set -x
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_h.log 2>&1
( time timeout 900 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_h.json 2> gpurun_out/bench_cfg3_h.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bscsr_(sample|replay)" -s 6 -c 2 -o gpurun_out/prof_bscsr_small_h python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_small_cfg3_h.log 2>&1
ls -la gpurun_out
