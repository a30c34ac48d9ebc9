"""CPU-side checks of the host logic: weight folding, scheduler mirrors, guidance slot tables, the C ABI's
symbol table, and the fail-loudly contract when no GPU is present.  No kernel runs here."""
import ctypes
import os
import math
import re
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.conditioning import BRANCH_STREAM, expand_guidance_batch, guidance_slots
from convofusion_b200.pack import STREAMS, fold_cross_attention
from oracle import convofusion_oracle as O
from helpers import SCHED_KW, rel_err, state_dict

ROOT = Path(__file__).resolve().parents[1]


def test_folded_cross_attention_is_exact_algebra():
    """pack.fold_cross_attention vs the as-written five MultiheadAttentions + att_fuser (cross_attention.py:578-652),
    evaluated in float64 so only algebra (not rounding) is compared."""
    sd = {k: v.double() for k, v in state_dict().items() if k.startswith("denoiser.decoder.layers.3.")}
    lp, d = "denoiser.decoder.layers.3.", 512
    g = torch.Generator().manual_seed(0)
    B, lens = 3, (5, 9, 7, 8, 1)
    t2 = torch.randn(16, B, d, generator=g, dtype=torch.float64)            # norm2(tgt)
    mems = [torch.randn(L, B, d, generator=g, dtype=torch.float64) * 2 + 0.3 for L in lens]
    masks = {"tlsn": torch.tensor([[0, 0, 0, 1, 1, 1, 1]] * B).bool(), "spkemb": torch.tensor([[0, 0, 1, 1, 1]] * B).bool()}
    outs = []
    for name, mem in zip(STREAMS, mems):
        m = O._ln(sd, lp + f"{name}_norm.", mem)
        o, _ = O._mha_sd(sd, lp + f"multihead_attn_{name}.", t2, m, m, 1, masks.get(name))
        outs.append(o)
    want = O._lin(sd, lp + "att_fuser.", torch.cat(outs, -1))
    w_qx, b_qx, w_fu, b_fu = fold_cross_attention(sd, lp, d)
    qx = F.linear(t2, w_qx, b_qx)                                           # [16,B,5d]
    us = []
    for x, (name, mem) in enumerate(zip(STREAMS, mems)):
        mu = mem.mean(-1, keepdim=True)
        xhat = (mem - mu) / torch.sqrt(((mem - mu) ** 2).mean(-1, keepdim=True) + 1e-5)
        s = torch.einsum("qbd,kbd->bqk", qx[..., x * d:(x + 1) * d], xhat)
        if masks.get(name) is not None:
            s = s.masked_fill(masks[name][:, None, :], float("-inf"))
        us.append(torch.einsum("bqk,kbd->qbd", torch.softmax(s, -1), xhat))
    got = F.linear(torch.cat(us, -1), w_fu, b_fu)
    assert rel_err(got, want) < 1e-12


@pytest.mark.parametrize("kind", ["ddim", "ddim_mld", "ddim_eta", "ddpm"])
def test_scheduler_mirror_matches_oracle(kind):
    """step_table() rows fed through the kernel's formula reproduce the oracle scheduler's step() exactly."""
    if kind == "ddpm":
        mine, ora = cf.DDPMScheduler(clip_sample=True, **SCHED_KW), O.DDPMSchedulerOracle(clip_sample=True, **SCHED_KW)
    elif kind == "ddim_mld":
        kw = dict(clip_sample=False, set_alpha_to_one=False, steps_offset=1)
        mine, ora = cf.DDIMScheduler(**kw, **SCHED_KW), O.DDIMSchedulerOracle(**kw, **SCHED_KW)
    else:
        mine, ora = cf.DDIMScheduler(clip_sample=True, **SCHED_KW), O.DDIMSchedulerOracle(clip_sample=True, **SCHED_KW)
    eta = 0.7 if kind == "ddim_eta" else 0.0
    n = 20
    tab = mine.step_table(n, eta=eta)
    ora.set_timesteps(n)
    assert tab["timesteps"].tolist() == ora.timesteps.tolist()
    assert tab["needs_noise"] == (kind in ("ddpm", "ddim_eta"))
    g = torch.Generator().manual_seed(4)
    x, eps, z = (torch.randn(2, 16, 128, generator=g) for _ in range(3))
    for i, t in enumerate(tab["timesteps"]):
        c = [torch.tensor(v) for v in tab["coef"][i]]
        x0 = (x - c[0] * eps) / c[1]
        if tab["clip_sample"]:
            x0 = x0.clamp(-1, 1)
        prev = c[2] * x0 + c[3] * (eps if kind != "ddpm" else x)
        if float(c[4]) != 0.0:
            prev = prev + c[4] * z
        kw = {"variance_noise": z} if kind in ("ddpm", "ddim_eta") else {}
        if kind.startswith("ddim"):
            kw["eta"] = eta
        want = ora.step(eps, int(t), x, **kw).prev_sample
        assert torch.equal(prev, want), (kind, i)
        a = ora.alphas_cumprod[int(t)]
        assert tab["coef"][i, 5] == np.float32(a ** 0.5) and tab["coef"][i, 6] == np.float32((1 - a) ** 0.5)


def test_scheduler_public_surface():
    import inspect
    d, p = cf.DDIMScheduler(**SCHED_KW), cf.DDPMScheduler(**SCHED_KW)
    assert "eta" in inspect.signature(d.step).parameters          # convofusion.py:427-429 detects DDIM this way
    assert "eta" not in inspect.signature(p.step).parameters
    for s in (d, p):
        assert s.init_noise_sigma == 1.0 and s.config.num_train_timesteps == 1000 and s.betas.shape == (1000,)
        s.set_timesteps(50)
        assert len(s.timesteps) == 50 and s.timesteps.dtype == torch.int64
    x0, n = torch.randn(2, 8, 128), torch.randn(2, 8, 128)
    assert torch.equal(p.add_noise(x0, n, torch.tensor(400)), O.DDPMSchedulerOracle(**SCHED_KW).add_noise(x0, n, 400))
    with pytest.raises(NotImplementedError):
        cf.DDIMScheduler(prediction_type="sample", **SCHED_KW)
    with pytest.raises(ValueError):
        d.set_timesteps(2000)


def test_guidance_slots_reproduce_the_seven_branch_batch():
    B = 3
    slots = guidance_slots(B, 7, "cpu")
    for x in range(5):
        s = slots[x].view(7, B)
        for g in range(7):
            cond = g == 6 or BRANCH_STREAM.get(g) == x
            assert s[g].tolist() == ([1, 2, 3] if cond else [0, 0, 0])
    assert all(len(s) == 6 * B for s in guidance_slots(B, 6, "cpu"))
    # monadic clips: the speaker-only branch (3) is left out, the remaining branches keep their order
    from convofusion_b200.conditioning import guidance_branches, guidance_slots_host
    assert guidance_branches() == [0, 1, 2, 3, 4, 5] and guidance_branches(True) == [0, 1, 2, 3, 4, 5, 6]
    br = guidance_branches(True, spk_is_uncond=True)
    assert br == [0, 1, 2, 4, 5, 6]
    mono = guidance_slots(B, br, "cpu")
    for x in range(5):
        assert mono[x].view(6, B).tolist() == [slots[x].view(7, B)[g].tolist() for g in br]
        assert torch.equal(mono[x], guidance_slots_host(B, br)[x])
    assert mono[0].view(6, B)[:5].sum() == 0      # no branch but the full one conditions on the speaker
    # gathering slots == torch.cat([...]*7) ordering of convofusion.py:911-929
    enc = [torch.arange(4.0).view(4, 1, 1).expand(4, 2, 3).contiguous() + 10 * x for x in range(5)]
    masks = {"tlsn": torch.tensor([[1, 1], [0, 1], [0, 0], [1, 0]]).bool(), "spkemb": None, "alsn": None}
    enc7, masks7 = expand_guidance_batch(enc, masks, B)
    u, c = torch.zeros(B), torch.arange(1.0, B + 1)
    want_tlsn = torch.cat([u, c, u, u, u, u, c]) + 20       # text_lsn list at :911
    want_alsn = torch.cat([u, u, c, u, u, u, c]) + 10       # melspec_lsn cat at :916
    want_spk = torch.cat([u, u, u, c, u, u, c])             # text_spk at :918
    want_apb = torch.cat([u, u, u, u, c, u, c]) + 30        # active_passive_bit at :921-927
    want_id = torch.cat([u, u, u, u, u, c, c]) + 40         # lsn_id at :929
    for got, want in zip(enc7, (want_spk, want_alsn, want_tlsn, want_apb, want_id)):
        assert got[:, 0, 0].tolist() == want.tolist()
    assert masks7["tlsn"].shape == (7 * B, 2) and masks7["tlsn"][B].tolist() == [False, True]
    assert masks7["alsn"] is None


def test_abi_exports_every_declared_symbol():
    header = (ROOT / "include" / "convofusion_b200.h").read_text()
    declared = set(re.findall(r"\b(cfb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cfb_abi_version() == 2


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_lib.DenoiserLayer) == 26 * 8
    assert ctypes.sizeof(_lib.DenoiserWeights) == 8 * 4 + 30 * 8
    assert ctypes.sizeof(_lib.Memory) == 20 * 8 + 10 * 4
    assert ctypes.sizeof(_lib.Schedule) == 16 + 16
    assert ctypes.sizeof(_lib.VaeLayer) == 20 * 8
    assert ctypes.sizeof(_lib.VaeDecoder) == 8 + 32 + 32 + 32 + 8
    assert ctypes.sizeof(_lib.VaeEncLayer) == 12 * 8
    assert ctypes.sizeof(_lib.VaeEncoder) == 8 + 32 + 32 + 5 * 8 + 8
    assert ctypes.sizeof(_lib.VaeWeights) == (24 + 16 + 2 * ctypes.sizeof(_lib.VaeDecoder) + 8 +
                                              2 * ctypes.sizeof(_lib.VaeEncoder))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks behaviour on a box without a GPU")
def test_no_cpu_fallback():
    den = cf.default_denoiser("fp32")
    with pytest.raises(_lib.CfbError):
        den(torch.zeros(7, 16, 128), torch.tensor(10), [torch.zeros(7, 4, 512)] * 5, None, {})
    with pytest.raises(_lib.CfbError):
        cf.default_vae("fp32").decode(torch.zeros(2, 1, 8, 128), [128])
    sch = cf.DDIMScheduler(**SCHED_KW)
    sch.set_timesteps(50)
    with pytest.raises(_lib.CfbError):
        sch.step(torch.zeros(1, 16, 128), 0, torch.zeros(1, 16, 128))
    w, h = _lib.DenoiserWeights(), ctypes.c_void_p()
    assert _lib.lib().cfb_denoiser_create(ctypes.byref(w), ctypes.byref(h)) == -3      # CFB_ERR_NO_DEVICE
    assert b"no CPU fallback" in _lib.lib().cfb_last_error()


def test_precision_switches_and_packed_formats():
    """The process-wide precision switches validate their arguments without a device, and the packer writes what the
    handles expect: bf16 + fp16 matrices for a 16-bit denoiser (cfb_denoiser_attach_f16_weights), fp16 or bf16 matrices
    for a 16-bit VAE depending on cfb_get_vae_f16, fp32 everywhere for fp32 handles."""
    from convofusion_b200.pack import pack_denoiser, pack_vae
    lib = _lib.lib()
    assert lib.cfb_set_bf16_activation_f16(32) != 0 and lib.cfb_set_bf16_activation_f16(-1) != 0
    assert lib.cfb_set_bf16_activation_sites(4) != 0 and lib.cfb_set_bf16_activation_sites(64) != 0
    assert lib.cfb_set_bf16_activation_terms(3) != 0
    for fn, good in ((lib.cfb_set_bf16_activation_f16, (0, 1, 19, 31)), (lib.cfb_set_bf16_activation_sites, (0, 2, 18, 27, 16))):
        for v in good:
            assert fn(v) == 0
    cf.set_bf16_activation_f16(True)
    cf.set_bf16_activation_sites(cf.ACT_SITE_LATENT_PROJ)
    sd = state_dict()
    den = {k[len("denoiser."):]: v for k, v in sd.items() if k.startswith("denoiser.")}
    vae = {k[len("vae."):]: v for k, v in sd.items() if k.startswith("vae.")}
    pk = pack_denoiser(den, "", 9, 4, 16, _lib.BF16, "cpu")
    assert {t.dtype for t in pk["keep"]} == {torch.bfloat16, torch.float32}
    assert {t.dtype for t in pk["keep16"]} == {torch.float16}
    w, w16 = pk["struct"], pk["struct16"]
    assert w16.n_layers == w.n_layers == 9 and w16.w_out and w16.w_embed and all(w16.w_zx[x] and w16.w_yx[x] for x in range(5))
    for l in range(9):
        for f in ("w_in", "w_so", "w_tb1", "w_tb2", "w_qx", "w_fu", "w_ff1", "w_ff2"):
            assert getattr(pk["layers16"][l], f) and getattr(pk["layers16"][l], f) != getattr(pk["layers"][l], f)
    # the fp16 matrices are roundings of the fp32 parameters, not of their bf16 forms
    w_in = den["decoder.layers.0.self_attn.in_proj_weight"]
    t16 = next(t for t in pk["keep16"] if t.shape == w_in.shape)
    assert torch.equal(t16, w_in.to(torch.float16)) and not torch.equal(t16, w_in.to(torch.bfloat16).to(torch.float16))
    assert "struct16" not in pack_denoiser(den, "", 9, 4, 16, _lib.F32, "cpu")
    try:
        for f16, want in ((1, torch.float16), (0, torch.bfloat16)):
            assert lib.cfb_set_vae_f16(f16) == 0 and lib.cfb_get_vae_f16() == f16
            mats = {t.dtype for t in pack_vae(vae, "", 5, 2, 1024, _lib.BF16, "cpu")["keep"]}
            assert mats == {want, torch.float32}
        assert {t.dtype for t in pack_vae(vae, "", 5, 2, 1024, _lib.F32, "cpu")["keep"]} == {torch.float32}
    finally:
        cf.set_vae_f16(True)


def test_lane_context_and_pool_host_logic():
    """modules.lane is a per-thread, re-entrant selector; SamplerPool validates its arguments and, like every other
    entry point, refuses to run without a GPU."""
    import threading
    from convofusion_b200.modules import current_lane, lane
    assert current_lane() == 0
    seen = {}
    with lane(2):
        assert current_lane() == 2
        with lane(1):
            assert current_lane() == 1
        assert current_lane() == 2
        t = threading.Thread(target=lambda: seen.setdefault("other", current_lane()))
        t.start(); t.join()
    assert current_lane() == 0 and seen["other"] == 0
    with pytest.raises(ValueError):
        cf.SamplerPool(None, lanes=0)
    assert cf.SamplerPool(None, lanes=2).chains == 3 and cf.SamplerPool(None, lanes=3).chains == 2   # measured defaults
    assert cf.SamplerPool(None, lanes=1).chains == 0 and cf.SamplerPool(None, lanes=4, chains=1).chains == 1
    if not torch.cuda.is_available():
        pool = cf.SamplerPool(cf.ConvoFusionSampler(precision="fp32"), lanes=2)
        with pytest.raises(_lib.CfbError):
            pool.generate_many([{}])


def test_modules_copy_without_their_device_handles():
    """deepcopy / pickle of a drop-in module must not duplicate the raw device handle (double free) nor trip over the
    ctypes structs of the packed weights: the copy starts unpacked."""
    import copy
    import pickle
    den = cf.default_denoiser("fp32")
    den._handle, den._lanes, den._packed = ctypes.c_void_p(0), {1: ctypes.c_void_p(0)}, {"struct": _lib.DenoiserWeights()}
    for clone in (copy.deepcopy(den), pickle.loads(pickle.dumps(den))):
        assert clone._handle is None and clone._lanes == {} and clone._packed is None
        assert torch.equal(clone.latent_embd.weight, den.latent_embd.weight)
    den._handle, den._lanes, den._packed = None, {}, None


def test_window_bookkeeping_matches_reference_process_text():
    """convofusion_b200.windows against strings produced by the reference's own process_text
    (unbounded_synthesis.py:189-241) on random word timings, stored by tools/pin_reference_loops.py."""
    from convofusion_b200.windows import slice_windows, window_spans, window_text
    from helpers import golden, unbounded_windows
    g = golden("ref_loops.pt")
    assert len(g["window_text_cases"]) >= 50
    for c in g["window_text_cases"]:
        assert window_text(c["segments"], c["t0"], c["t1"]) == c["text"]
    assert window_text("-" * 10, 0.0, 5.12) == "-" * 10
    assert window_spans(256) == [(0.0, 5.12), (2.56, 7.68), (5.12, 10.24)] and len(window_spans(58 * 128)) == 115
    wins, _, _ = unbounded_windows(g["unbounded"], g["B"])       # also asserts the window texts of the synthetic batch
    assert [w["mel_lsn"].shape[1] for w in wins] == [161] * 3 and [w["apb"].shape[1] for w in wins] == [8] * 3


def test_state_dict_layout_is_the_reference_layout():
    """Spot-check the key convention of SURVEY 8b (full strict-load against the reference modules is done by
    tools/make_golden.py, which needs /root/reference)."""
    keys = set(state_dict())
    for k in ("denoiser.cond_params", "denoiser.query_pos.pe", "denoiser.decoder.layers.8.multihead_attn_lsnemb.in_proj_weight",
              "denoiser.decoder.layers.0.time_block2.emb_layers.1.bias", "denoiser.decoder.layers.4.time_block1.out_layers.2.weight",
              "denoiser.decoder.layers.2.apb_norm.weight", "denoiser.decoder.norm.bias", "vae.body_decoder.linear_blocks.1.weight",
              "vae.hands_encoder.input_blocks.0.self_attn.out_proj.bias", "vae.body_global_motion_token",
              "vae.mem_pos_decoder.pe", "text_audio_encoder.text_encoder.projection.1.weight",
              "text_audio_encoder.audio_encoder.main.3.bias", "text_audio_encoder.audio_time_proj.weight",
              "condition_fuser.lsn_id_emb.weight", "condition_fuser.latent_proj.2.bias"):
        assert k in keys, k
    assert state_dict()["text_audio_encoder.audio_time_proj.weight"].shape == (512, 161)
    assert sum(v.numel() for k, v in state_dict().items() if k.startswith("denoiser.") and not k.endswith(".pe")) == 92923013


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) runs without a GPU and prints one JSON line with
    the contract's keys, on our arm's metric / unit / workload."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ddim-steps", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "motion_seconds_per_second" and line["unit"] == "motion-s/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "motion-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["clips_per_gpu"] == 64 and "configs[1]" in line["config"]["workload"]


def test_bench_line_assembly_carries_the_contract_keys():
    """bench.assemble_line (pure host arithmetic over the measured times) for the default, single-lane, unbounded and
    multi-GPU cases: contract keys present, value / e2e / roofline arithmetic consistent."""
    import importlib.util
    import types
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("bench_under_test", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    parts = {"conditioning_ms": 0.9, "loop_ms": 64.0, "decode_ms": 0.9}
    clocks = {"sm_mhz": 1965, "sm_max_mhz": 1965, "reasons": [], "samples": 4}
    for windows, world, F in ((0, 1, 2), (0, 1, 1), (12, 1, 1), (0, 8, 2)):
        args = types.SimpleNamespace(steps=16, warmup=3, ddim_steps=50, precision="bf16", dyadic=False, windows=windows,
                                     no_cpu_baseline=True, batch=64, sweep=0)
        single = {"value": 5000.0, "unit": "motion-s/s", "ms_per_step": 65.5} if F > 1 else None
        per_step = 53.7 * max(1, windows)                               # a step of W serial windows takes W passes
        line = bench.assemble_line(args, world=world, B=64, F=F, n_branch=6, ms_dev=16 * per_step, ms_e2e=16 * per_step * 1.01,
                                   launches=400000, clocks=clocks, parts=parts,
                                   roof={"tflops": 451.0, "ms": 0.74, "launches": 72, "kernel": "k", "mem": {"bound": "hbm"}},
                                   single=single,
                                   h2d_bytes=16441344, d2h_bytes=6193152)
        json_line = __import__("json").dumps(line)                      # serialisable
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert key in line, key
        assert line["n_gpus"] == world and line["dtype"] == "bf16" and line["vs_baseline"] is None and json_line
        assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
        motion_s = 5.12 if windows == 0 else 5.12 * (windows + 1) / 2
        assert abs(line["value"] - 64 * world * motion_s / (per_step * 1e-3)) < 1e-6 * line["value"]
        assert abs(line["ms_per_step"] - per_step) < 1e-9 and line["execution"]["batches_in_flight"] == F
        assert line["roofline_mem"] == {"bound": "hbm"} and line["execution"]["guidance_branches_evaluated"] == 6
        r = line["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert r["peak"] >= r["sustained_peak"] > 0 and (r["traffic"] is None or r["traffic"] > 0)
        assert 0 < r["whole_step_frac"] < 1 and 0 < r["whole_step_frac_at_throughput"] < 1
        assert "workload" in line["config"] and "model" not in line["config"]
