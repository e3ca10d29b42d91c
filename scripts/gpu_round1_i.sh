# round 1, GPU call i: scalar-list replay kernel (parity + time), 896/1024-thread prefetch variants of the stream kernel
mkdir -p gpurun_out
set -x
( time timeout 1200 python -m pytest tests/test_gpu_bscsr.py tests/test_gpu_golden.py tests/test_gpu_host_exe.py -x -q ) > gpurun_out/pytest_bscsr_i.log 2>&1
for v in 0 36 35; do
  ( TKS_BSCSR_VARIANT=$v timeout 600 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_i_v$v.json 2> gpurun_out/bench_cfg3_i_v$v.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg3_i.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3_i.log 2>&1
ls -la gpurun_out
