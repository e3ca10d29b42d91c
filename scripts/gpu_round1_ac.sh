# round 1, GPU call ac: timing experiment -- what does the row-start bitmap stream (32 B per warp iteration) cost?
# TKS_DBG_ROWBITS_STRIDE=0 pins the bitmap pointer (results are wrong, the kernel's other work is unchanged)
mkdir -p gpurun_out
for st in 32 0 32 0; do
  TKS_DBG_ROWBITS_STRIDE=$st timeout 300 python - <<PY >> gpurun_out/ac_summary.txt 2>&1
import sys, numpy as np
sys.path.insert(0, ".")
from _pkg import pkg
tks = pkg()
eng = tks.SpMV(num_cols=1024, k=100, profile_kernels=True)
eng.generate_synthetic(10_000_000, 1024, 20, "gamma", seed=0)
rng = np.random.default_rng(1)
ms = []
for i in range(40):
    v = rng.random(1024); v = (v / np.linalg.norm(v)).astype(np.float32)
    eng.reset(v); eng.run_timed(100)
    if i >= 10: ms.append(eng.stats().last_main_kernel_ms)
print("stride $st main_kernel_ms", float(np.mean(ms)), "candidates", eng.stats().last_candidates)
PY
done
cat gpurun_out/ac_summary.txt
