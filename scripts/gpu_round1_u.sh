# round 1, GPU call u (2 GPUs): programmatic dependent launch of the per-query kernels, peer-memory candidate exchange
# (one kernel: store to every rank's IPC window over NVLink + wait + merge) against the NCCL all-gather path
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_csr.py tests/test_gpu_golden.py tests/test_gpu_batched.py -x -q ) > gpurun_out/pytest_gpu_u.log 2>&1
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_u_n1.json 2> gpurun_out/bench_cfg2_u_n1.err
( TKS_PDL=0 timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_u_n1_nopdl.json 2> gpurun_out/bench_cfg2_u_n1_nopdl.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/bench_cfg2_u_n2.json 2> gpurun_out/bench_cfg2_u_n2.err
( TKS_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/bench_cfg2_u_n2_nccl.json 2> gpurun_out/bench_cfg2_u_n2_nccl.err
ls -la gpurun_out
