"""bf16 handles with two-term activations (`cfb_set_bf16_activation_sites(mask)`): the LayerNorm outputs feeding the GEMM
sites in the mask (1 qkv, 2 TimeBlock linears, 8 linear1, 16 latent_proj) are stored as hi + lo bf16 terms and those
GEMMs issue two accumulating tcgen05.mma per K step (gemm_tc.cu A2); the weights stay bf16.  tools/precision_study.py
and tools/split_sites.py predict the gains (activation rounding is what the -36.5 / +7.5 guidance weights amplify);
checked here against the oracle and the reference golden.  Mask 16 is the library default."""
import pytest
import torch

import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.conditioning import expand_guidance_batch
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import SCHED_KW, golden, oracle_batch, oracle_denoise, rel_err, state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DEFAULT = (31, 16)   # library defaults: (cfb_set_bf16_activation_f16, cfb_set_bf16_activation_sites)


def sampler(steps):
    s = cf.ConvoFusionSampler(precision="bf16")
    s.load_state_dict(state_dict())
    s = s.to(DEV).eval()
    s.scheduler = cf.DDIMScheduler(clip_sample=True, **SCHED_KW)
    s.num_inference_timesteps = steps
    return s


def test_two_term_activations_cut_the_bf16_error():
    s = sampler(50)
    lib = _lib.lib()
    # one evaluation, dyadic B = 3 with ragged masks: general per-pair path, all operators
    syn = synthetic_clip(3, seed=78, dyadic=True)
    d = to_device(syn, DEV)
    enc, masks = s.encode_conditions(d["clip"], d["uncond_text"], d["uncond_text_attn"])
    enc7, masks7 = expand_guidance_batch(enc, masks, 3)
    x = torch.randn(21, 16, 128, generator=torch.Generator().manual_seed(6))
    want, _ = oracle_denoise(x, 500, *oracle_batch(syn))
    # the benchmarked structure (shared-slot plan, chains, graph): B = 1 DDIM-50 golden of the reference modules
    g = golden("sample_ddim50_clip.pt")
    syn1 = synthetic_clip(1, seed=1235, dyadic=False)
    d1 = to_device(syn1, DEV)
    enc_s, masks_s = s.encode_conditions(d1["clip"], d1["uncond_text"], d1["uncond_text_attn"])
    init = torch.randn(1, 16, 128, generator=torch.Generator().manual_seed(100)).to(DEV)
    eps_default, _ = s.denoiser(x.to(DEV), torch.tensor(500), enc7, None, masks7)
    e, l = {}, {}
    try:
        for f16, mask in ((0, 0), (1, 0), (3, 0), (7, 0), (15, 0), (31, 0), (0, 16), (31, 16), (0, 27)):
            _lib.check(lib.cfb_set_bf16_activation_f16(f16))
            _lib.check(lib.cfb_set_bf16_activation_sites(mask))
            eps, att = s.denoiser(x.to(DEV), torch.tensor(500), enc7, None, masks7)
            _, rec, _ = s.sample(enc_s, masks_s, 1, init, record=True)
            _, rec_eager, _ = s.sample(enc_s, masks_s, 1, init, record=True, use_graph=False)
            assert torch.equal(rec, rec_eager), (f16, mask)       # graph replay == eager launches in every mode
            for i in range(5):
                assert float(att[i].sum(-1).sub(1).abs().max()) < 1e-4
            k = (f16, mask)
            e[k] = rel_err(eps.cpu(), want)
            l[k] = [rel_err(rec[i].cpu(), g["record"][i]) for i in (0, 24, 49)]
            print(f"fp16 operand groups {f16:2d}, two-term sites {mask:2d}: one evaluation eps L2 vs oracle "
                  f"{e[k]:.2e}; DDIM-50 latents L2 vs reference golden after steps 1 / 25 / 50: "
                  f"{l[k][0]:.3f} {l[k][1]:.3f} {l[k][2]:.3f}")
            if k == DEFAULT:
                assert torch.equal(eps, eps_default)              # the library default; switching modes is stateless
        assert lib.cfb_set_bf16_activation_sites(4) != 0          # scores / conditional queries are not a two-term site
    finally:
        _lib.check(lib.cfb_set_bf16_activation_f16(DEFAULT[0]))
        _lib.check(lib.cfb_set_bf16_activation_sites(DEFAULT[1]))
    plain = l[(0, 0)]
    assert e[(0, 27)] < e[(0, 0)]   # a single evaluation is dominated by the (branch-common) weight rounding: small gain
    assert l[(0, 16)][2] < 0.7 * plain[2] and l[(0, 16)][0] < 0.7 * plain[0]   # latent_proj alone
    assert l[(0, 27)][0] < 0.5 * plain[0] and l[(0, 27)][2] < 0.5 * plain[2] and l[(0, 27)][2] < 0.1
    # fp16 LayerNorm outputs: every LayerNorm-fed site at 11 bits for free; all operand groups: a quarter of the error
    assert l[(1, 0)][0] < 0.5 * plain[0] and l[(1, 0)][2] < 0.5 * plain[2] and l[(1, 0)][2] < 0.1
    assert l[(31, 0)][0] < 0.3 * plain[0] and l[(31, 0)][2] < 0.3 * plain[2]
    _lib.check(lib.cfb_set_bf16_activation_terms(2))              # shorthand: every site
    _lib.check(lib.cfb_set_bf16_activation_f16(0))
    try:
        _, rec27, _ = s.sample(enc_s, masks_s, 1, init, record=True)
    finally:
        _lib.check(lib.cfb_set_bf16_activation_f16(DEFAULT[0]))
        _lib.check(lib.cfb_set_bf16_activation_sites(DEFAULT[1]))
    assert rel_err(rec27[49].cpu(), g["record"][49]) == l[(0, 27)][2]
