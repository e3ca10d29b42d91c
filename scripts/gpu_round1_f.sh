# round 1, GPU call f: tail chunks + L2 prefetch variants of the BS-CSR stream kernel, full suite, cfg2 and cfg3 lines
mkdir -p gpurun_out
set -x
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_f.log 2>&1
for v in 0 34 16 33; do
  ( TKS_BSCSR_VARIANT=$v timeout 600 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_f_v$v.json 2> gpurun_out/bench_cfg3_f_v$v.err
done
( time timeout 600 python bench.py ) > gpurun_out/bench_cfg2_f.json 2> gpurun_out/bench_cfg2_f.err
ls -la gpurun_out
