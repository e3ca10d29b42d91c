# round 2, GPU call ao (2 GPUs): pipelined partition exchange of the FPGA mode (ShardedSpMVFixed.submit / fetch) -- the
# 2-rank NCCL test, then cfg3 over 2 GPUs through torchrun (value = pipelined form, `blocking` beside it)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -3
( timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29981 bench.py --gpus 2 --workload cfg3 --steps 20 --warmup 5 ) > gpurun_out/r02ao_bench_cfg3_n2.json 2> gpurun_out/r02ao_bench_cfg3_n2.err
python - <<'PY'
import json
try:
    j=json.loads(open("gpurun_out/r02ao_bench_cfg3_n2.json").read().strip().splitlines()[-1])
    print("cfg3 n2", round(j["ms_per_step"],4), j["value"], "blocking", j["blocking"], "local", j["local_kernels_ms"], j["parity_n"], j["parity"])
except Exception as e: print("ERR", e, open("gpurun_out/r02ao_bench_cfg3_n2.err").read()[-2500:])
PY
