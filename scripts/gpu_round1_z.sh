# round 1, GPU call z (8 GPUs): the BASELINE multi-GPU configurations -- weak-scaled cfg2 (peer exchange and NCCL),
# cfg4 (200M x 1024 uniform-40 row-sharded over 8), cfg5 (64 queries x 50M over 8)
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_z.txt
run() { name=$1; shift; ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 "$@" ) > gpurun_out/bench_z_$name.json 2> gpurun_out/bench_z_$name.err; PORT=$((PORT+1)); }
PORT=29520
run cfg2_n8_peer --steps 50 --warmup 5
TKS_EXCHANGE=nccl run cfg2_n8_nccl --steps 50 --warmup 5
run cfg4_n8 --workload cfg4 --steps 20 --warmup 3
run cfg5_n8 --workload cfg5 --steps 5 --warmup 3
ls -la gpurun_out | tail -12
