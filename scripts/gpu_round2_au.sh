# round 2, GPU call au (1 GPU): ncu evidence of the final code -- launch list of the default bench command, full captures
# of the float main kernel as it runs by default (12-bit column offsets; fp32 and half), the batched main kernel and the
# BS-CSR stream kernel; profiles/traffic.json is rewritten from them
mkdir -p gpurun_out
GIT=$1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02au_cfg2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-sub > gpurun_out/r02au_launches.log 2>&1
cap() { name=$1; kern=$2; skip=$3; shift; shift; shift; timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$kern -s $skip -c 2 -f -o gpurun_out/r02au_$name "$@" > gpurun_out/r02au_$name.log 2>&1; }
cap cfg2_csr_topk_main_kernel csr_topk_main_kernelILi256ELi0 5 python bench.py --steps 3 --warmup 3 --no-cpu --no-sub
python scripts/ncu_summary.py gpurun_out/r02au_cfg2_csr_topk_main_kernel.ncu-rep gpurun_out/r02au_cfg2_csr_topk_main_kernel --traffic-key cfg2 --kernel csr_topk_main_kernel --git $GIT | head -30
cap cfg2h_csr_topk_main_kernel csr_topk_main_kernelILi256ELi1 5 python bench.py --workload cfg2h --steps 3 --warmup 3 --no-cpu
python scripts/ncu_summary.py gpurun_out/r02au_cfg2h_csr_topk_main_kernel.ncu-rep gpurun_out/r02au_cfg2h_csr_topk_main_kernel --traffic-key cfg2h --kernel csr_topk_main_kernel --git $GIT | head -30
cap cfg5_csr_batched_kernel csr_batched_kernelILb0ELb0 2 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu
python scripts/ncu_summary.py gpurun_out/r02au_cfg5_csr_batched_kernel.ncu-rep gpurun_out/r02au_cfg5_csr_batched_kernel --traffic-key cfg5 --kernel csr_batched_kernel --git $GIT | head -30
cap cfg3_bscsr_stream_kernel bscsr_stream_kernel 3 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-uniform
python scripts/ncu_summary.py gpurun_out/r02au_cfg3_bscsr_stream_kernel.ncu-rep gpurun_out/r02au_cfg3_bscsr_stream_kernel --traffic-key cfg3 --kernel bscsr_stream_kernel --git $GIT | head -30
cp profiles/traffic.json gpurun_out/r02au_traffic.json
rm -f gpurun_out/r02au_*.ncu-rep
ls -la gpurun_out/r02au_* | head -30
