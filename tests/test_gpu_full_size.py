"""Parity at BASELINE.json's FULL size (10M x 1024 gamma-20, 1.95e8 non-zeros): scripts/full_size_parity.py runs the
oracle's literal sequential kernel over the whole matrix (about 7 s per mode on the host) and compares
  cfg3: every result word of every partition and the merged list, bit for bit -- reference semantics and drift-free
        mode, host-packed and GPU-packed packets;
  cfg2: the fp32 top-100 against the reference gold restatement (identical index set, scores within 1e-5 relative);
  cfg2h: the same in the half-precision value mode against the gold on half-rounded inputs."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_full_size_cfg2_cfg3_parity(cuda_required):
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "full_size_parity.py"), "--rows", "10000000"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout)
    assert res["nnz"] > 1.9e8 and res["all_ok"]
    assert res["cfg2"]["same_index_set"] and res["cfg2"]["max_rel_score_diff"] < 1e-5
    assert res["cfg2h"]["same_index_set"] and res["cfg2h"]["max_rel_score_diff"] < 1e-5   # 16-bit values, fp32 sums
    assert len(res["cfg3"]) == 4
    for name, r in res["cfg3"].items():
        assert r["result_words_identical"] and r["merged_list_identical"], name
