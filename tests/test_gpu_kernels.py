"""Per-kernel checks on the B200 through the C ABI (cfb_linear / cfb_layernorm / cfb_mha /
cfb_guidance_sched_step / cfb_audio_encoder) against the oracle's torch-CPU arithmetic."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import convofusion_b200 as cf
from convofusion_b200 import _lib
from oracle import convofusion_oracle as O
from helpers import SCHED_KW, max_rel, rel_err, state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def linear(A, W, bias=None, act="none", a_act="none", out_bf16=False, accumulate_into=None, backend=_lib.GEMM_AUTO):
    a_bf16 = A.dtype == torch.bfloat16
    M, K = A.shape
    N = W.shape[0]
    out = accumulate_into if accumulate_into is not None else torch.empty(
        M, N, device=A.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    _lib.check(_lib.lib().cfb_linear(A.data_ptr(), int(a_bf16), W.data_ptr(), _lib.ptr(bias), out.data_ptr(),
                                     int(out_bf16), M, N, K, _lib.ACT[act], _lib.ACT[a_act],
                                     int(accumulate_into is not None), backend, _lib.stream_ptr()))
    return out


@pytest.mark.parametrize("M,N,K", [(1, 512, 512), (50, 2048, 512), (96, 512, 128), (777, 1536, 512), (322, 256, 80),
                                   (128, 69, 128), (130, 120, 1024)])
def test_simt_gemm_fp32(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    got = linear(A.to(DEV), W.to(DEV), b.to(DEV), act="gelu", backend=_lib.GEMM_SIMT).cpu()
    assert max_rel(got, F.gelu(F.linear(A, W, b))) < 2e-6
    got = linear(A.to(DEV), W.to(DEV), b.to(DEV), a_act="silu", backend=_lib.GEMM_SIMT).cpu()
    assert max_rel(got, F.linear(F.silu(A), W, b)) < 2e-6
    base = torch.randn(M, N, generator=g)
    acc = base.clone().to(DEV)
    linear(A.to(DEV), W.to(DEV), b.to(DEV), accumulate_into=acc, backend=_lib.GEMM_SIMT)
    assert max_rel(acc.cpu(), base + F.linear(A, W, b)) < 2e-6


TC_SHAPES = [(128, 128, 64), (96, 512, 512), (6144, 512, 512), (777, 1536, 512), (1536, 2560, 512), (300, 512, 2560),
             (1000, 1024, 512), (515, 512, 1024), (256, 128, 512), (2048, 384, 128), (64, 64, 128), (200, 32, 64),
             (129, 96, 192)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_tcgen05_gemm_matches_simt_and_fp64(M, N, K):
    """The tcgen05/TMA kernel against (a) the CUDA-core kernel on the same bf16 operands and (b) float64 matmul of
    the bf16-rounded operands: fp32 accumulation of exact bf16 products leaves only summation-order noise."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).bfloat16()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g)
    ref = A.double() @ W.double().T + b.double()
    Ad, Wd, bd = A.to(DEV), W.to(DEV), b.to(DEV)
    tc = linear(Ad, Wd, bd, backend=_lib.GEMM_TCGEN05).cpu()
    simt = linear(Ad, Wd, bd, backend=_lib.GEMM_SIMT).cpu()
    assert max_rel(tc, ref) < 5e-6, "tcgen05 vs float64"
    assert max_rel(tc, simt) < 5e-6
    # bf16 output + GELU epilogue
    tcb = linear(Ad, Wd, bd, act="gelu", out_bf16=True, backend=_lib.GEMM_TCGEN05).float().cpu()
    assert max_rel(tcb, F.gelu(ref.float())) < 6e-3
    # residual accumulate
    base = torch.randn(M, N, generator=g)
    acc = base.clone().to(DEV)
    linear(Ad, Wd, bd, accumulate_into=acc, backend=_lib.GEMM_TCGEN05)
    assert max_rel(acc.cpu(), base.double() + ref) < 5e-6


def test_tcgen05_is_the_default_bf16_engine():
    before = _lib.lib().cfb_launch_count()
    A = torch.randn(256, 512, device=DEV).bfloat16()
    W = torch.randn(512, 512, device=DEV).bfloat16()
    out = linear(A, W)
    assert _lib.lib().cfb_launch_count() == before + 1
    assert max_rel(out.cpu(), A.double().cpu() @ W.double().cpu().T) < 5e-6
    with pytest.raises(ValueError):   # K=80 cannot take the tensor-core path and the caller forced it
        linear(torch.zeros(8, 80, device=DEV).bfloat16(), torch.zeros(32, 80, device=DEV).bfloat16(), backend=_lib.GEMM_TCGEN05)


@pytest.mark.parametrize("d,rows", [(512, 1), (512, 1000), (128, 777)])
def test_layernorm(d, rows):
    g = torch.Generator().manual_seed(d + rows)
    x = torch.randn(rows, d, generator=g) * 3 + 1
    w, b = torch.randn(d, generator=g), torch.randn(d, generator=g)
    out = torch.empty(rows, d, device=DEV)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)     # keep device operands alive across the raw-pointer call
    _lib.check(_lib.lib().cfb_layernorm(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), 0, rows, d,
                                        _lib.stream_ptr()))
    assert max_rel(out.cpu(), F.layer_norm(x, (d,), w, b)) < 2e-6
    outb = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16)
    _lib.check(_lib.lib().cfb_layernorm(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), outb.data_ptr(), 1, rows, d,
                                        _lib.stream_ptr()))
    assert max_rel(outb.float().cpu(), F.layer_norm(x, (d,), w, b)) < 5e-3


@pytest.mark.parametrize("n,Lq,Lk,H,hd,ragged", [(14, 16, 16, 4, 128, False), (5, 128, 128, 2, 64, True), (5, 128, 8, 2, 64, False),
                                                 (3, 37, 37, 2, 64, True)])
def test_mha_core(n, Lq, Lk, H, hd, ragged):
    """cfb_mha vs the oracle's restatement of nn.MultiheadAttention with identity projections."""
    E = H * hd
    g = torch.Generator().manual_seed(n + Lq + Lk)
    q, k, v = (torch.randn(L, n, E, generator=g) for L in (Lq, Lk, Lk))
    lens = torch.randint(1, Lk + 1, (n,), generator=g) if ragged else None
    kpm = (torch.arange(Lk)[None] >= lens[:, None]) if ragged else None
    eye = torch.eye(E)
    want, _ = O.mha(q, k, v, torch.cat([eye, eye, eye]), torch.zeros(3 * E), eye, torch.zeros(E), H, kpm)
    qd, kd, vd = (t.permute(1, 0, 2).contiguous().to(DEV) for t in (q, k, v))      # sample-major rows
    out = torch.empty(n * Lq, E, device=DEV)
    lens_d = lens.int().to(DEV) if ragged else None
    _lib.check(_lib.lib().cfb_mha(qd.data_ptr(), E, kd.data_ptr(), vd.data_ptr(), E, out.data_ptr(), E, 0, n, Lq, Lk, H, hd,
                                  _lib.ptr(lens_d), _lib.stream_ptr()))
    assert max_rel(out.view(n, Lq, E).permute(1, 0, 2).cpu(), want) < 5e-6


@pytest.mark.parametrize("n,Lq,Lk,H,hd,ragged", [(14, 16, 16, 4, 128, False), (5, 128, 128, 2, 64, True), (5, 128, 8, 2, 64, False),
                                                 (24, 18, 18, 2, 64, True), (3, 37, 37, 2, 64, True), (2, 100, 77, 2, 64, True)])
def test_mha_core_bf16_tensor_core(n, Lq, Lk, H, hd, ragged):
    """bf16 cfb_mha (mma.sync kernel: scores in registers, probabilities rounded to bf16 for P.V) vs the fp32 oracle
    on the same bf16-rounded inputs.  Tolerance: two bf16 roundings (P and the output), 2^-8 of the output scale."""
    E = H * hd
    g = torch.Generator().manual_seed(100 + n + Lq + Lk)
    q, k, v = (torch.randn(L, n, E, generator=g).bfloat16().float() for L in (Lq, Lk, Lk))
    lens = torch.randint(1, Lk + 1, (n,), generator=g) if ragged else None
    kpm = (torch.arange(Lk)[None] >= lens[:, None]) if ragged else None
    eye = torch.eye(E)
    want, _ = O.mha(q, k, v, torch.cat([eye, eye, eye]), torch.zeros(3 * E), eye, torch.zeros(E), H, kpm)
    qd, kd, vd = (t.permute(1, 0, 2).contiguous().bfloat16().to(DEV) for t in (q, k, v))
    out = torch.full((n * Lq, E), float("nan"), device=DEV, dtype=torch.bfloat16)
    lens_d = lens.int().to(DEV) if ragged else None
    _lib.check(_lib.lib().cfb_mha(qd.data_ptr(), E, kd.data_ptr(), vd.data_ptr(), E, out.data_ptr(), E, 1, n, Lq, Lk, H, hd,
                                  _lib.ptr(lens_d), _lib.stream_ptr()))
    got = out.float().view(n, Lq, E).permute(1, 0, 2).cpu()
    assert torch.isfinite(got).all()
    assert max_rel(got, want) < 8e-3


@pytest.mark.parametrize("kind", ["ddim", "ddpm"])
def test_guidance_scheduler_step_is_bit_exact(kind):
    """Fused combine + step vs the oracle (torch eager fp32, same association order): identical bits."""
    B, n = 5, 16 * 128
    g = torch.Generator().manual_seed(11)
    eps7 = torch.randn(7 * B, 16, 128, generator=g)
    x, z = torch.randn(B, 16, 128, generator=g), torch.randn(B, 16, 128, generator=g)
    if kind == "ddim":
        mine, ora = cf.DDIMScheduler(clip_sample=True, **SCHED_KW), O.DDIMSchedulerOracle(clip_sample=True, **SCHED_KW)
    else:
        mine, ora = cf.DDPMScheduler(clip_sample=True, **SCHED_KW), O.DDPMSchedulerOracle(clip_sample=True, **SCHED_KW)
    tab = mine.step_table(10, eta=0.0)
    ora.set_timesteps(10)
    for i in (0, 4, 9):
        t = int(tab["timesteps"][i])
        kw = {"variance_noise": z} if kind == "ddpm" else {"eta": 0.0}
        want = ora.step(O.guidance_combine(eps7, 7.5), t, x, **kw).prev_sample
        coef = torch.tensor(tab["coef"][i]).to(DEV)
        zd = z.to(DEV)
        for nb in (7, 6):   # 6 = the weight-0 full-cond branch is not evaluated at all
            xd = x.clone().to(DEV)
            ed = eps7[: nb * B].contiguous().to(DEV)
            _lib.check(_lib.lib().cfb_guidance_sched_step(ed.data_ptr(), xd.data_ptr(), zd.data_ptr(), coef.data_ptr(), nb, B,
                                                          n, tab["kind"], 1, 7.5, _lib.stream_ptr()))
            assert torch.equal(xd.cpu(), want), (kind, i, nb)
    # the scheduler mirror's own step() (single-branch path of the same kernel)
    mine.set_timesteps(10)
    e1 = O.guidance_combine(eps7, 7.5)
    kw = {"variance_noise": z.to(DEV)} if kind == "ddpm" else {"eta": 0.0}
    got = mine.step(e1.to(DEV), 500, x.to(DEV), **kw).prev_sample
    kw = {"variance_noise": z} if kind == "ddpm" else {"eta": 0.0}
    assert torch.equal(got.cpu(), ora.step(e1, 500, x, **kw).prev_sample)


def test_conditioning_projections():
    s = cf.ConvoFusionSampler(precision="fp32")
    s.load_state_dict(state_dict())
    s = s.to(DEV)
    g = torch.Generator().manual_seed(7)
    mel = torch.rand(2, 24, 80, generator=g) * 80 - 80
    t5 = torch.randn(2, 6, 768, generator=g)
    sd = state_dict()
    assert max_rel(s.text_audio_encoder.audio_encoder(mel.to(DEV)).cpu(), O.audio_encoder(sd, mel)) < 5e-6
    txt, _, _ = s.text_audio_encoder.text_encoder(t5.to(DEV), torch.ones(2, 6, device=DEV))
    assert max_rel(txt.cpu(), O.text_projection(sd, t5)) < 5e-6


def test_keypoints3d_is_bit_exact():
    """base.py:204-209 post-processing (divide by 3, re-attach fingers to wrists, everything to the root)."""
    x = torch.randn(5, 37, 189, generator=torch.Generator().manual_seed(8)) * 2
    got = cf.keypoints3d(x.to(DEV))
    want = O.feats_to_keypoints3d(x.reshape(-1, 189)).reshape(5, 37, 63, 3)
    assert got.shape == (5, 37, 63, 3) and torch.equal(got.cpu(), want)
    with pytest.raises(_lib.CfbError):
        cf.keypoints3d(x)


def test_motion_writer_npy_layout(tmp_path):
    """pred.npy / gt.npy / att_*.npy as base.py:save_npy lays them out (key points [length, 63, 3], cut to the
    sample's length; attention map of batch entry 0 per recorded timestep), written behind an asynchronous D2H copy."""
    import numpy as np
    g = torch.Generator().manual_seed(21)
    pred, gt = torch.randn(3, 128, 189, generator=g), torch.randn(3, 128, 189, generator=g)
    lengths, keys = [128, 100, 37], ["a_1", "b_2", "c_3"]
    att = {981: [torch.rand(3, 9, 16, m, generator=g) for m in (32, 161, 32, 8, 1)],
           1: [torch.rand(3, 9, 16, m, generator=g) for m in (32, 161, 32, 8, 1)]}
    with cf.MotionWriter(tmp_path, ring=2) as w:
        for rep in range(3):          # more batches than ring slots: slots are recycled
            w.submit(pred.to(DEV) + rep, lengths, [k + f"_{rep}" for k in keys], m_ref=gt.to(DEV),
                     att_maps={t: [a.to(DEV) for a in maps] for t, maps in att.items()} if rep == 0 else None)
        w.flush()
        assert w.files_written == 3 * 2 * 3 + 3 * 2 * 5
    for rep in range(3):
        want = O.feats_to_keypoints3d((pred + rep).reshape(-1, 189)).reshape(3, 128, 63, 3).numpy()
        want_gt = O.feats_to_keypoints3d(gt.reshape(-1, 189)).reshape(3, 128, 63, 3).numpy()
        for i, (k, L) in enumerate(zip(keys, lengths)):
            p = np.load(tmp_path / f"{k}_{rep}" / "pred.npy")
            assert p.shape == (L, 63, 3) and p.dtype == np.float32 and np.array_equal(p, want[i, :L])
            assert np.array_equal(np.load(tmp_path / f"{k}_{rep}" / "gt.npy"), want_gt[i, :L])
    for k in keys:
        for name, maps_t in (("att_alsn", 1), ("att_lsnemb", 4)):
            a = np.load(tmp_path / f"{k}_0" / name / "att_981.npy")
            assert np.array_equal(a, att[981][maps_t][0].numpy())
    with pytest.raises(_lib.CfbError):
        cf.MotionWriter(tmp_path).submit(pred, lengths, keys)
