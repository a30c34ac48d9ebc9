"""The opt-in cta_group::2 (CTA pair, 256x256 tile) tcgen05 GEMM, `CFB_TC_2CTA=1`.  Kept in its own, last-sorting
file so nothing else queues behind these tests; every run is a subprocess under a timeout (round 1 saw one hang with
two batches in flight: two pair CTAs per SM deadlocking on the tensor-memory allocation permits, fixed by keeping one
pair CTA per SM)."""
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


LANES_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import convofusion_b200 as cf
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import state_dict
s = cf.ConvoFusionSampler(precision="bf16", num_inference_timesteps=4)
s.load_state_dict(state_dict()); s = s.to("cuda:0").eval()
batches = []
for k in range(6):
    syn = to_device(synthetic_clip(16, seed=50 + k, dyadic=bool(k %% 2)), "cuda:0")
    init = torch.randn(16, 16, 128, generator=torch.Generator().manual_seed(60 + k)).cuda()
    batches.append(dict(clip=syn["clip"], uncond_text=syn["uncond_text"], uncond_text_attn=syn["uncond_text_attn"],
                        lengths=[128] * 16, init_latents=init))
pool = cf.SamplerPool(s, lanes=3)
outs = pool.generate_many(batches)
torch.cuda.synchronize()
torch.save(torch.stack([o["lat_t"].cpu() for o in outs]), sys.argv[1])
""" % (str(ROOT), str(ROOT / "tests"))


def test_cta_pair_with_three_batches_in_flight(tmp_path):
    """The round-1 hang scenario: several host threads / streams launching graphs full of CTA-pair GEMMs next to
    single-CTA ones.  Must finish (timeout) and agree with the default kernels."""
    res = {}
    for name, env in (("base", {}), ("pair", {"CFB_TC_2CTA": "1"})):
        out = tmp_path / f"{name}.pt"
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", LANES_SCRIPT, str(out)], env=e, capture_output=True, text=True, timeout=240)
        assert r.returncode == 0, r.stderr[-2000:]
        res[name] = torch.load(out, weights_only=True)
    l2 = float((res["pair"] - res["base"]).norm() / res["base"].norm())
    print(f"three lanes, CTA-pair GEMMs vs default: latents after 4 steps L2 {l2:.2e}")
    assert torch.isfinite(res["pair"]).all() and l2 < 0.2
