# round 1, GPU call v (2 GPUs): where do the ~36 us of the 2-GPU step go?  No exchange at all vs peer kernel vs NCCL
mkdir -p gpurun_out
set -x
for mode in none peer nccl none peer; do
  ( TKS_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 ) > gpurun_out/bench_cfg2_v_n2_$mode.json 2> gpurun_out/bench_cfg2_v_n2_$mode.err
  python -c "
import json;j=json.loads(open('gpurun_out/bench_cfg2_v_n2_$mode.json').read().strip().splitlines()[-1]);print('$mode',j['ms_per_step'],j['e2e']['ms_per_step'])" >> gpurun_out/v_summary.txt
done
( timeout 600 python bench.py --no-cpu --steps 50 ) > gpurun_out/bench_cfg2_v_n1.json 2> gpurun_out/bench_cfg2_v_n1.err
cat gpurun_out/v_summary.txt
