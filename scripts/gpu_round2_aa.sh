# round 2, GPU call aa (1 GPU): batched kernel with every quad's batch starting on a 128-byte line -- parity, cfg5, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batched.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02aa_bench_cfg5.json 2> gpurun_out/r02aa_bench_cfg5.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02aa_bench_cfg5.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("cfg5", round(j["ms_per_step"],4), "main_alone", round(r["main_kernel_ms"],4), "lds", r["lds"]["frac"], "e2e", round(j["e2e"]["ms_per_step"],4), j["parity_n"])
PY
timeout 600 ncu --set full --clock-control none --kernel-name-base mangled -k regex:csr_batched_kernelILb0ELb0 -s 2 -c 1 -f -o gpurun_out/r02aa_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02aa_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02aa_cfg5.ncu-rep gpurun_out/r02aa_cfg5_csr_batched_kernel | head -20
rm -f gpurun_out/r02aa_cfg5.ncu-rep
