# round 1, GPU call d: BSX packet format -- BS-CSR parity tests, cfg3 variant sweep, ncu of the new stream kernel
mkdir -p gpurun_out
set -x
( time timeout 1200 python -m pytest tests/test_gpu_bscsr.py tests/test_gpu_golden.py -x -q ) > gpurun_out/pytest_bscsr_d.log 2>&1
for v in 0 33 16 1; do
  ( TKS_BSCSR_VARIANT=$v timeout 600 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_d_v$v.json 2> gpurun_out/bench_cfg3_d_v$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bscsr_stream -s 3 -c 1 -o gpurun_out/prof_bscsr_stream_d python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg3_d.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg3_d.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3_d.log 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_bscsr.py --deselect tests/test_gpu_golden.py ) > gpurun_out/pytest_gpu_d.log 2>&1
ls -la gpurun_out
