"""Optional execution strategies (read from the environment when the library initialises, hence subprocesses):
staged ld/st GEMM epilogue, strided softmax, CUDA-core attention, single-chain / no-PDL execution, the general
per-pair path instead of the shared-slot plan.  Each must reproduce the default configuration's bf16 sampling run
(same arithmetic up to summation order), and the default run must be bit-identical across processes."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import convofusion_b200 as cf
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import state_dict
s = cf.ConvoFusionSampler(precision="bf16", num_inference_timesteps=2)
s.load_state_dict(state_dict()); s = s.to("cuda:0").eval()
syn = to_device(synthetic_clip(8, seed=41, dyadic=True), "cuda:0")
enc, masks = s.encode_conditions(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"])
init = torch.randn(8, 16, 128, generator=torch.Generator().manual_seed(42)).cuda()
z, rec, _ = s.sample(enc, masks, 8, init, record=True)
torch.save(rec.cpu(), sys.argv[1])
""" % (str(ROOT), str(ROOT / "tests"))


def run(tmp_path, name, env):
    out = tmp_path / f"{name}.pt"
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", SCRIPT, str(out)], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out, weights_only=True)


def test_optional_paths_agree_with_default(tmp_path):
    base = run(tmp_path, "base", {})
    again = run(tmp_path, "again", {})
    assert torch.equal(base, again)                      # deterministic across processes
    scale = float(base[0].abs().max())
    # the two-term latent_proj operand (default, CFB_BF16_ACT_SITES=16) exists for the TMA-epilogue GEMM only: the staged
    # epilogue is compared with plain operands on both sides
    plain = run(tmp_path, "plain", {"CFB_BF16_ACT_SITES": "0"})
    for name, env in (("staged_epilogue", {"CFB_TC_TMA_EPI": "0", "CFB_BF16_ACT_SITES": "0"}),
                      ("softmax_strided", {"CFB_SOFTMAX_STRIDED": "1"}), ("mha_simt", {"CFB_MHA_SIMT": "1"}),
                      ("no_plan", {"CFB_PLAN": "0"}), ("rowblock_all", {"CFB_ROWBLOCK": "7"}),
                      ("cross_tcgen05", {"CFB_CROSS_TC": "1"}),
                      ("serial", {"CFB_CHAINS": "1", "CFB_PDL": "0", "CFB_OVERLAP": "0"})):
        got = run(tmp_path, name, env)
        err = float((got[0] - base[0]).abs().max()) / scale
        l2 = float((got[0] - base[0]).norm() / base[0].norm())
        print(f"{name}: first-step deviation from default: max {err:.2e}, L2 {l2:.2e}")
        # chains / PDL change no arithmetic at all; the TMA store / reduce-add epilogue uses the same arithmetic as the
        # staged one; the row-block kernel and the general per-pair path change summation order and a few rounding
        # sites, which flips bf16 roundings of GEMM operands (amplified ~74x by the guidance weights)
        if name in ("serial", "staged_epilogue", "softmax_strided"):
            assert torch.equal(got, plain if name == "staged_epilogue" else base), name
        else:
            # mha_simt: CUDA-core attention keeps the probabilities in fp32, the tensor-core kernel rounds them to bf16
            # before P.V; no_plan: the general per-pair path rounds the projected queries instead of the pre-projected
            # keys / values; rowblock: other summation order + one-pass LayerNorm statistics.  Each moves bf16 rounding
            # sites, amplified like every other one.  Measured with fp16 activations (the default): mha_simt 0.006,
            # no_plan 0.008, cross_tcgen05 0.017 (bf16 queries / memory in that kernel), rowblock_all 0.037 (the row-block
            # programs keep bf16 activations)
            assert l2 < 0.08, name

