# round 1, GPU call g: drift-free fixed mode (parity vs its oracle, recall), full suite, cfg3 line with recall of both modes
mkdir -p gpurun_out
set -x
( time timeout 1200 python -m pytest tests/test_gpu_bscsr.py -x -q ) > gpurun_out/pytest_bscsr_g.log 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_bscsr.py ) > gpurun_out/pytest_gpu_g.log 2>&1
( time timeout 900 python bench.py --workload cfg3 --steps 20 ) > gpurun_out/bench_cfg3_g.json 2> gpurun_out/bench_cfg3_g.err
ls -la gpurun_out
