# round 1, GPU call m: 32-copy query variant of the float main kernel: parity, A/B against the single-copy kernel, ncu
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests/test_gpu_csr.py tests/test_gpu_golden.py tests/test_gpu_batched.py tests/test_gpu_multirank.py -x -q ) > gpurun_out/pytest_gpu_m.log 2>&1
( timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_m_xrep32.json 2> gpurun_out/bench_cfg2_m_xrep32.err
( TKS_CSR_XREP=1 timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_m_xrep1.json 2> gpurun_out/bench_cfg2_m_xrep1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_topk_main -s 3 -c 1 -o gpurun_out/prof_csr_main_m python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg2_m.log 2>&1
ls -la gpurun_out
