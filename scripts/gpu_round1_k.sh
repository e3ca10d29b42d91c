# round 1, GPU call k (2 GPUs): the driver's own sequence -- smoke, default bench at N=2 (weak-scaled cfg2) and its reference arm
mkdir -p gpurun_out
set -x
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/smoke_k.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/bench_n2_k.json 2> gpurun_out/bench_n2_k.err
( time timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/bench_n1_k.json 2> gpurun_out/bench_n1_k.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_ref_n2_k.json 2> gpurun_out/bench_ref_n2_k.err
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q ) > gpurun_out/pytest_multirank_k.log 2>&1
ls -la gpurun_out
