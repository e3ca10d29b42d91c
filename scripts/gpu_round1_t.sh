# round 1, GPU call t (2 GPUs): multi-rank tests over NCCL (float shards, FPGA-mode partitions over ranks), weak-scaled
# cfg2 at N=1 and N=2, reference arm at N=2
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_t.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_t.log 2>&1
( time timeout 600 python bench.py --no-cpu ) > gpurun_out/bench_cfg2_t_n1.json 2> gpurun_out/bench_cfg2_t_n1.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/bench_cfg2_t_n2.json 2> gpurun_out/bench_cfg2_t_n2.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg4 --steps 10 --warmup 3 ) > gpurun_out/bench_cfg4_t_n2.json 2> gpurun_out/bench_cfg4_t_n2.err
ls -la gpurun_out
