# round 2, GPU call bc (1 GPU): the C++ host's throughput loop (-F n: submit / fetch on the float, half, fixed and group
# functors, checked against the blocking verbs by the executable itself)
timeout 200 python -m pytest tests/test_gpu_host_exe.py -x -q 2>&1 | tail -4
