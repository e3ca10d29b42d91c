# round 1, GPU call ap: bfloat16 value mode (tests, bench cfg2b next to cfg2h / cfg2), full-size parity test inside pytest
mkdir -p gpurun_out
set -x
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_ap.log 2>&1
tail -4 gpurun_out/pytest_gpu_ap.log
for wl in cfg2b cfg2h cfg2; do
  ( timeout 600 python bench.py --no-cpu --workload $wl ) > gpurun_out/bench_${wl}_ap.json 2> gpurun_out/bench_${wl}_ap.err
  python -c "
import json;j=json.loads(open('gpurun_out/bench_${wl}_ap.json').read().strip().splitlines()[-1]);print('$wl',j['ms_per_step'],j['roofline']['main_kernel_ms'],j['e2e']['ms_per_step'],j['roofline']['streamed_frac'])"
done
