# round 2, GPU call n (1 GPU): single-process multi-GPU group API (two..eight shards of one device), -G list of the C++
# host; compute-sanitizer over every kernel family; whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_host_exe.py -x -q 2>&1 | tail -8
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
bash scripts/sanitize.sh gpurun_out
tail -5 gpurun_out/sanitize_memcheck.log
