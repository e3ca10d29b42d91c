# round 1, GPU call c: full GPU suite after the BS-CSR stream rework, stream-kernel variant sweep, cfg2/cfg5 re-measure
mkdir -p gpurun_out
set -x
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_c.log 2>&1
for v in 0 33 16 8 1; do
  ( TKS_BSCSR_VARIANT=$v timeout 600 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_v$v.json 2> gpurun_out/bench_cfg3_v$v.err
done
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_c.json 2> gpurun_out/bench_cfg2_c.err
( time timeout 900 python bench.py --workload cfg5 --steps 10 --no-cpu ) > gpurun_out/bench_cfg5_c.json 2> gpurun_out/bench_cfg5_c.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bscsr_stream -s 3 -c 1 -o gpurun_out/prof_bscsr_stream_c python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg3_c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg3_c.csv python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg3_c.log 2>&1
ls -la gpurun_out
