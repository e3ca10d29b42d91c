# round 1, GPU call e (2 GPUs): one process per GPU over NCCL -- multi-rank parity test, cfg4 and cfg5 benches at N=2
mkdir -p gpurun_out
set -x
nvidia-smi -L > gpurun_out/smi_e.txt
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q ) > gpurun_out/pytest_multirank_e.log 2>&1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/bench_cfg4_n2.json 2> gpurun_out/bench_cfg4_n2.err
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5 --steps 10 --warmup 3 ) > gpurun_out/bench_cfg5_n2.json 2> gpurun_out/bench_cfg5_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
ls -la gpurun_out
