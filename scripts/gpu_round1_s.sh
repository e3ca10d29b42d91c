# round 1, GPU call s: GPU-side BS-CSR packer (tests + cfg3 bench with pack timings), tau back in the sample kernel's
# last block with the one-pass histogram threshold, cheaper chunk-edge masks
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s.log 2>&1
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_s.json 2> gpurun_out/bench_cfg2_s.err
( time timeout 900 python bench.py --workload cfg3 --steps 20 --no-cpu ) > gpurun_out/bench_cfg3_s.json 2> gpurun_out/bench_cfg3_s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2_s.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_cfg2_s.log 2>&1
ls -la gpurun_out
