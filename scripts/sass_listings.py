"""Regenerate profiles/r02_sass_*.txt: `cuobjdump -sass` of the kernels that dominate the four modes, as they are built
into approximate-spmv-topk_b200/lib/libtopkspmv.so (CPU only; the listings show LDG.E.256 / LDS.128 / SHFL / REDUX and, for
the bulk-copy variant, UBLKCP / SYNCS).

    python scripts/sass_listings.py"""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "approximate-spmv-topk_b200" / "lib" / "libtopkspmv.so"
WANT = {   # output name -> mangled-name fragment
    "csr_topk_main_kernel_256_0_c12": "csr_topk_main_kernelILi256ELi0ELb0ELb1ELi0E",
    "csr_topk_main_kernel_256_1_c12": "csr_topk_main_kernelILi256ELi1ELb0ELb1ELi0E",
    "csr_sample_kernel_0_c12": "csr_sample_kernelILi0ELb1E",
    "csr_batched_kernel_0_0": "csr_batched_kernelILb0ELb0E",
    "bscsr_stream_kernel_20_4_32_768_1_1": "bscsr_stream_kernelILi20ELi4ELi32ELi768ELb1ELb1E",
    "select_topk_kernel_1": "select_topk_kernelILb1E",
}
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"(?m)^(?=\s*Function : )", sass)
for name, frag in WANT.items():
    hit = [b for b in blocks if b.lstrip().startswith("Function : ") and frag in b.splitlines()[0]]
    assert len(hit) == 1, (name, len(hit))
    out = ROOT / "profiles" / f"r02_sass_{name}.txt"
    out.write_text(hit[0])
    ops = re.findall(r"(?m)^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", hit[0])
    print(f"{out.name}: {len(ops)} instructions, LDG {sum(o.startswith('LDG') for o in ops)}, LDS {sum(o.startswith('LDS') for o in ops)}, "
          f"SHFL {sum(o.startswith('SHFL') for o in ops)}")
