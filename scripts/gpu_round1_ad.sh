# round 1, GPU call ad: new CLI / profile-switch tests with the full suite
mkdir -p gpurun_out
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_ad.log 2>&1
tail -5 gpurun_out/pytest_gpu_ad.log
