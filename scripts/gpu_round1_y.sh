# round 1, GPU call y (2 GPUs): multirank tests incl. k = 300 / 1024 through the peer exchange
mkdir -p gpurun_out
set -x
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q ) > gpurun_out/pytest_gpu_y.log 2>&1
tail -5 gpurun_out/pytest_gpu_y.log
