# round 1, GPU call o: float main kernel with two iterations in flight (2 CTAs x 448 threads) vs one (2 x 512)
mkdir -p gpurun_out
set -x
( TKS_CSR_DEPTH=2 timeout 1500 python -m pytest tests/test_gpu_csr.py tests/test_gpu_golden.py -x -q ) > gpurun_out/pytest_gpu_o_depth2.log 2>&1
( timeout 1500 python -m pytest tests/test_gpu_csr.py tests/test_gpu_golden.py tests/test_gpu_multirank.py -x -q ) > gpurun_out/pytest_gpu_o.log 2>&1
for d in 1 2; do
  ( TKS_CSR_DEPTH=$d timeout 600 python bench.py --no-cpu --steps 30 ) > gpurun_out/bench_cfg2_o_depth$d.json 2> gpurun_out/bench_cfg2_o_depth$d.err
done
TKS_CSR_DEPTH=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_topk_main -s 3 -c 1 -o gpurun_out/prof_csr_main_o python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_cfg2_o.log 2>&1
ls -la gpurun_out
