"""The opt-in cta_group::2 (CTA pair, 256x256 tile) tcgen05 GEMM, `CFB_TC_2CTA=1`.  Kept in its own, last-sorting
file: the kernel is verified single-stream only, so nothing else queues behind these tests."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from test_gpu_options import run

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_cta_pair_gemm_kernel():
    """tcgen05.mma.cta_group::2 GEMM against float64 on shapes with partial row tiles, an odd number of 128-row tiles,
    K = 64..1024 and all three epilogues (tools/pair_check.py)."""
    e = dict(os.environ)
    e["CFB_TC_2CTA"] = "1"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "pair_check.py"), "check"], env=e, capture_output=True,
                       text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_cta_pair_sampling_run_agrees_with_default(tmp_path):
    """A bf16 sampling run with every N % 256 == 0 GEMM on the CTA-pair kernel (one chain, no side streams) against the
    default configuration: same bf16 products, fp32 accumulation inside the tensor core in its own order."""
    base = run(tmp_path, "base", {})
    got = run(tmp_path, "cta_pair", {"CFB_TC_2CTA": "1", "CFB_CHAINS": "1", "CFB_OVERLAP": "0"})
    l2 = float((got[0] - base[0]).norm() / base[0].norm())
    print(f"cta_pair: first-step deviation from default: L2 {l2:.2e}, identical: {torch.equal(got, base)}")
    assert l2 < 1e-2
