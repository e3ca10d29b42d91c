# round 2, GPU call u (1 GPU): ncu of the bulk-copy variant of the 16-bit main kernel (why it is slower), its pytest
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tma.py -x -q 2>&1 | tail -3
TKS_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:csr_topk_main_kernelILi256ELi1ELb1 -s 5 -c 1 -f -o gpurun_out/r02u_cfg2h_tma python bench.py --workload cfg2h --steps 3 --warmup 3 --no-cpu > gpurun_out/r02u_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02u_cfg2h_tma.ncu-rep gpurun_out/r02u_cfg2h_csr_topk_main_kernel_tma | head -30
rm -f gpurun_out/r02u_cfg2h_tma.ncu-rep
