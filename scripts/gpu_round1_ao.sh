# round 1, GPU call ao (2 GPUs): exchange fused into the select kernel -- multirank tests, probe, bench peer vs nccl
mkdir -p gpurun_out
set -x
( time timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_csr.py -x -q ) > gpurun_out/pytest_gpu_ao.log 2>&1
tail -4 gpurun_out/pytest_gpu_ao.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/peer_exchange_probe.py > gpurun_out/probe_ao.log 2>&1
grep rank gpurun_out/probe_ao.log | cut -c1-300
for mode in peer nccl peer; do
  ( TKS_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 ) > gpurun_out/bench_cfg2_ao_n2_$mode.json 2> gpurun_out/bench_cfg2_ao_n2_$mode.err
  python -c "
import json;j=json.loads(open('gpurun_out/bench_cfg2_ao_n2_$mode.json').read().strip().splitlines()[-1]);print('$mode',j['ms_per_step'],j['e2e']['ms_per_step'],j['gpu_launches'])"
done
