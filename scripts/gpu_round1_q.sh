# round 1, GPU call q: half-precision value mode (tests + cfg2h bench + ncu), same-box comparison with the
# reference's own GPU host (cuSPARSE / LightSpMV + thrust sort)
mkdir -p gpurun_out
set -x
( time timeout 1200 python -m pytest tests/test_gpu_csr.py tests/test_gpu_batched.py tests/test_gpu_host_exe.py -x -q ) > gpurun_out/pytest_gpu_q.log 2>&1
( time timeout 600 python bench.py --workload cfg2h --no-cpu ) > gpurun_out/bench_cfg2h_q.json 2> gpurun_out/bench_cfg2h_q.err
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_q.json 2> gpurun_out/bench_cfg2_q.err
( time timeout 1500 python scripts/compare_reference_gpu.py --rows 2000000 --iters 12 ) > gpurun_out/compare_ref_gpu_q.json 2> gpurun_out/compare_ref_gpu_q.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_topk_main -s 3 -c 1 -o gpurun_out/prof_csr_main_half_q python bench.py --workload cfg2h --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg2h_q.log 2>&1
ls -la gpurun_out
