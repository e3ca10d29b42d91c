# round 1, GPU call x (2 GPUs): low-latency peer exchange (tagged 16-byte records, no fence / flag): tests, probe, bench
mkdir -p gpurun_out
set -x
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q ) > gpurun_out/pytest_gpu_x.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/peer_exchange_probe.py > gpurun_out/probe_x.log 2>&1
for mode in peer nccl none; do
  ( TKS_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 ) > gpurun_out/bench_cfg2_x_n2_$mode.json 2> gpurun_out/bench_cfg2_x_n2_$mode.err
done
grep rank gpurun_out/probe_x.log
