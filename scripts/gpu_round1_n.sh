# round 1, GPU call n: float main kernel -- two-deep chunk scheduler, L2 prefetch distance sweep, 1 vs 32 query copies
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests/test_gpu_csr.py tests/test_gpu_golden.py tests/test_gpu_multirank.py -x -q ) > gpurun_out/pytest_gpu_n.log 2>&1
for cfg in "0 1" "3 1" "6 1" "0 32" "3 32" "6 32"; do
  set -- $cfg
  ( TKS_CSR_L2PF=$1 TKS_CSR_XREP=$2 timeout 600 python bench.py --no-cpu --steps 30 ) > gpurun_out/bench_cfg2_n_pf$1_x$2.json 2> gpurun_out/bench_cfg2_n_pf$1_x$2.err
done
ls -la gpurun_out
