# round 2, GPU call ae (2 GPUs): FPGA mode over several GPUs with the result words all-gathered device to device --
# bit-exactness test, cfg3 over 2 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -k fixed 2>&1 | tail -3
( timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29950 bench.py --gpus 2 --workload cfg3 --steps 20 --warmup 5 ) > gpurun_out/r02ae_bench_cfg3_n2.json 2> gpurun_out/r02ae_bench_cfg3_n2.err
python - <<'PY'
import json
try:
    j=json.loads(open("gpurun_out/r02ae_bench_cfg3_n2.json").read().strip().splitlines()[-1])
    print(j["ms_per_step"], j["value"], j["local_kernels_ms"], j["parity_n"], j["parity"])
except Exception as e: print("ERR", e, open("gpurun_out/r02ae_bench_cfg3_n2.err").read()[-1500:])
PY
