"""fp32 handles with their GEMMs on the tcgen05 tensor cores (three-way bf16 split of both operands, csrc/gemm_split.cu,
cfb_set_fp32_tensor_cores): the same goldens and the same 1e-4 tolerances as the CUDA-core fp32 mode, so the parity tests
exercise the TMA / tcgen05 / tensor-memory path at fp32 accuracy."""
import time

import pytest
import torch

import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.conditioning import expand_guidance_batch
from convofusion_b200.synthetic import synthetic_clip, to_device
from oracle import convofusion_oracle as O
from helpers import SCHED_KW, frac_within, golden, max_rel, oracle_batch, oracle_denoise, rel_err, state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture()
def fp32_tc():
    _lib.check(_lib.lib().cfb_set_fp32_tensor_cores(1))
    yield
    _lib.check(_lib.lib().cfb_set_fp32_tensor_cores(1))     # the library default


def sampler(steps, scheduler=None):
    s = cf.ConvoFusionSampler(precision="fp32")
    s.load_state_dict(state_dict())
    s = s.to(DEV).eval()
    s.scheduler = scheduler if scheduler is not None else cf.DDIMScheduler(clip_sample=True, **SCHED_KW)
    s.num_inference_timesteps = steps
    return s


def test_denoiser_forward_fp32_tc_vs_reference(fp32_tc):
    s = sampler(50)
    g = golden("denoiser_dyad_b2.pt")
    syn = synthetic_clip(2, seed=1236, dyadic=True)
    d = to_device(syn, DEV)
    enc, masks = s.encode_conditions(d["clip"], d["uncond_text"], d["uncond_text_attn"])
    enc7, masks7 = expand_guidance_batch(enc, masks, 2)
    x = torch.randn(2, 16, 128, generator=torch.Generator().manual_seed(101))
    n0 = _lib.lib().cfb_launch_count()
    eps, att = s.denoiser(sample=torch.cat([x] * 7).to(DEV), timestep=torch.tensor(g["t"], device=DEV),
                          encoder_hidden_states=enc7, lengths=None, mem_mask_dict=masks7)
    err = max_rel(eps.cpu(), g["eps"])
    _lib.check(_lib.lib().cfb_set_fp32_tensor_cores(0))
    eps_cc, _ = s.denoiser(sample=torch.cat([x] * 7).to(DEV), timestep=torch.tensor(g["t"], device=DEV),
                           encoder_hidden_states=enc7, lengths=None, mem_mask_dict=masks7)
    _lib.check(_lib.lib().cfb_set_fp32_tensor_cores(1))
    err_cc = max_rel(eps_cc.cpu(), g["eps"])
    print(f"fp32 eps max-rel vs reference: tensor cores (3 x bf16 split) {err:.2e}, CUDA cores {err_cc:.2e}; "
          f"launches {_lib.lib().cfb_launch_count() - n0}")
    assert err < 1e-4
    assert not torch.equal(eps, eps_cc)        # the switch selects another engine
    for a, ga in zip(att, g["att_full"]):
        assert max_rel(a.chunk(7)[-1].cpu(), ga) < 1e-3


def test_sampling_run_fp32_tc_vs_reference(fp32_tc):
    """BASELINE.json configs[0] (B = 1, DDIM-50, guidance 7.5, decode) on the split tensor-core GEMMs: every step
    within the fp32 tolerance of the reference's golden, graph replay == eager launches."""
    s = sampler(50)
    g = golden("sample_ddim50_clip.pt")
    syn = synthetic_clip(1, seed=1235, dyadic=False)
    d = to_device(syn, DEV)
    enc, masks = s.encode_conditions(d["clip"], d["uncond_text"], d["uncond_text_attn"])
    init = torch.randn(1, 16, 128, generator=torch.Generator().manual_seed(100)).to(DEV)
    z, rec, att = s.sample(enc, masks, 1, init, record=True, return_attention=True, use_graph=True)
    rec = rec.cpu()
    fr = [frac_within(rec[i], g["record"][i], 1e-4) for i in range(50)]
    l2 = [rel_err(rec[i], g["record"][i]) for i in range(50)]
    print(f"fp32-tc: min frac within 1e-4: {min(fr):.4f}; L2 rel first/last: {l2[0]:.2e}/{l2[-1]:.2e}")
    assert min(fr) >= 0.9
    assert max(l2) < 2e-4
    joints = s.decode(z, [128])
    assert max_rel(joints.cpu(), g["joints"]) < 1e-3
    assert max_rel(att[2][-1].cpu(), g["att_last_tlsn"]) < 2e-3
    z2, rec2, _ = s.sample(enc, masks, 1, init, record=True, use_graph=False)
    assert torch.equal(rec2.cpu(), rec)


def test_shared_slot_plan_fp32_tc_vs_oracle_and_speed(fp32_tc):
    """The benchmarked structure (shared-slot plan, chains, graph) with split tensor-core GEMMs against the oracle at
    1e-4, dyadic B = 5, three steps; then the throughput of the mode against the CUDA-core fp32 mode at batch 16."""
    sf = sampler(3)
    syn = synthetic_clip(5, seed=909, dyadic=True)
    init = torch.randn(5, 16, 128, generator=torch.Generator().manual_seed(910))
    d = to_device(syn, DEV)
    enc, masks = sf.encode_conditions(d["clip"], d["uncond_text"], d["uncond_text_attn"])
    enc_o, masks_o = oracle_batch(syn)
    want = []
    O.diffusion_reverse(oracle_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED_KW), enc_o, masks_o, init, 3,
                        record=want)
    _, rec, _ = sf.sample(enc, masks, 5, init.to(DEV), record=True)
    fr = [frac_within(rec[i].cpu(), want[i], 1e-4) for i in range(3)]
    l2 = [rel_err(rec[i].cpu(), want[i]) for i in range(3)]
    print(f"fp32-tc plan vs oracle: frac within 1e-4 {min(fr):.4f}, L2 {l2[0]:.2e}..{l2[-1]:.2e}")
    assert min(fr) >= 0.9 and max(l2) < 2e-4

    sf.num_inference_timesteps = 10
    syn = to_device(synthetic_clip(16, seed=5, dyadic=False), DEV)
    enc, masks = sf.encode_conditions(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"])
    init = torch.randn(16, 16, 128, generator=torch.Generator().manual_seed(3)).to(DEV)
    times = {}
    for mode in (1, 0):
        _lib.check(_lib.lib().cfb_set_fp32_tensor_cores(mode))
        sf.sample(enc, masks, 16, init)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sf.sample(enc, masks, 16, init)
        torch.cuda.synchronize()
        times[mode] = (time.perf_counter() - t0) / 10
    _lib.check(_lib.lib().cfb_set_fp32_tensor_cores(1))
    print(f"fp32 denoiser step at batch 16: tensor cores {times[1] * 1e3:.2f} ms, CUDA cores {times[0] * 1e3:.2f} ms "
          f"({times[0] / times[1]:.1f}x)")
    assert times[1] < times[0]
