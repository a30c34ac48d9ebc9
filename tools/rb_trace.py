#!/usr/bin/env python
"""Phase timeline of one row-block CTA (csrc/rowblock.cu, globaltimer stamps of the first row block of one launch):
where the time of a program goes -- residual load, operand latency, MMA issue, accumulator completion, statistics
pass, LayerNorm pass, stores.  `python tools/rb_trace.py [B] [chains]` prints the three program kinds of layer 4."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import state_dict

NAMES = ["A/residual ready", "first operands", "MMAs issued", "accumulators done", "stats pass done", "LN pass done", "stores issued"]


def main(B=64, chains=1):
    dev = "cuda:0"
    lib = _lib.lib()
    s = cf.ConvoFusionSampler(precision="bf16", num_inference_timesteps=1)
    s.load_state_dict(state_dict())
    s = s.to(dev).eval()
    s.denoiser.step_chains = chains
    _lib.check(lib.cfb_set_rowblock(7))
    syn = to_device(synthetic_clip(B, seed=3, dyadic=False), dev)
    init = torch.randn(B, 16, 128, generator=torch.Generator().manual_seed(4)).to(dev)
    enc, masks = s.encode_conditions(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"])
    s.sample(enc, masks, B, init, spk_is_uncond=True, use_graph=False)       # warm
    torch.cuda.synchronize()
    for kind, label in ((0, "out_proj -> TimeBlock1 -> norm2"), (1, "fuser/values -> TimeBlock2 -> norm3"), (2, "linear2 -> norm1")):
        lib.cfb_debug_rb_trace_arm(kind, 4, 1)
        s.sample(enc, masks, B, init, spk_is_uncond=True, use_graph=False)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 64)()
        lib.cfb_debug_rb_trace_read(out)
        t = list(out)
        t0 = t[0]
        print(f"program {kind} ({label}), B={B}, chains={chains}: total {((t[63] - t0) / 1e3):.1f} us")
        for i in range(7):
            row = t[1 + 8 * i: 1 + 8 * i + 7]
            if not any(row):
                continue
            print(f"  stage {i}: " + ", ".join(f"{n} {((v - t0) / 1e3):.1f}" for n, v in zip(NAMES, row) if v))
    lib.cfb_set_rowblock(0)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
