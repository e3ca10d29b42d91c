# round 2, GPU call k (1 GPU): quad batched kernel without the register swap -- parity tests, cfg5 on one GPU, ncu of it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batched.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02k_bench_cfg5.json 2> gpurun_out/r02k_bench_cfg5.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02k_bench_cfg5.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("cfg5", round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "lds", r.get("lds"), "e2e", round(j["e2e"]["ms_per_step"],4))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_batched_kernel -s 2 -c 1 -o gpurun_out/r02k_cfg5_batched python bench.py --workload cfg5 --rows 6000000 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02k_ncu.log 2>&1
ncu -i gpurun_out/r02k_cfg5_batched.ncu-rep --page details > gpurun_out/r02k_cfg5_csr_batched_kernel_ncu_details.txt 2>&1
grep -n -i "Duration\|Issue Slots Busy\|Eligible Warps\|Executed Ipc Active\|Registers Per\|bank conflicts\|Stall" gpurun_out/r02k_cfg5_csr_batched_kernel_ncu_details.txt | head -20
