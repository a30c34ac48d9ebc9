#!/usr/bin/env python
"""Correctness + timing of the cta_group::2 (CTA pair, 256x256 tile) tcgen05 GEMM against float64 and against the
128x128 single-CTA kernel.  The mode is read from the environment at library load:
    CFB_TC_2CTA=1 python tools/pair_check.py check     # exits non-zero on a mismatch
    CFB_TC_2CTA={0,1} python tools/pair_check.py time   # us / TFLOP/s per shape, graph-replayed launches"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from convofusion_b200 import _lib

dev = "cuda:0"
lib = _lib.lib()


def linear(A, W, b, out, act=0, acc=0, st=None):
    st = st if st is not None else torch.cuda.current_stream().cuda_stream
    _lib.check(lib.cfb_linear(A.data_ptr(), 1, W.data_ptr(), _lib.ptr(b), out.data_ptr(), int(out.dtype == torch.bfloat16),
                              A.shape[0], W.shape[0], A.shape[1], act, 0, acc, _lib.GEMM_TCGEN05, st))


def check():
    bad = 0
    for (M, N, K) in ((256, 256, 64), (256, 256, 512), (128, 256, 512), (1000, 512, 512), (6144, 512, 512), (777, 1536, 512),
                      (515, 512, 1024), (1536, 2560 - 256 * 2, 448), (4097, 1024, 512), (300, 4608, 512)):
        g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
        A = torch.randn(M, K, generator=g).bfloat16()
        W = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
        b = torch.randn(N, generator=g)
        ref = A.double() @ W.double().T + b.double()
        Ad, Wd, bd = A.to(dev), W.to(dev), b.to(dev)
        out = torch.empty(M, N, device=dev)
        linear(Ad, Wd, bd, out)
        torch.cuda.synchronize()
        e1 = float((out.cpu().double() - ref).abs().max() / ref.abs().max())
        base = torch.randn(M, N, generator=g)
        acc = base.clone().to(dev)
        linear(Ad, Wd, bd, acc, acc=1)
        torch.cuda.synchronize()
        e2 = float((acc.cpu().double() - (base.double() + ref)).abs().max() / ref.abs().max())
        ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        linear(Ad, Wd, bd, ob, act=_lib.ACT["gelu"])
        torch.cuda.synchronize()
        rg = torch.nn.functional.gelu(ref.float())
        e3 = float((ob.float().cpu() - rg).abs().max() / rg.abs().max())
        ok = e1 < 5e-6 and e2 < 5e-6 and e3 < 6e-3
        bad += not ok
        print(f"{M:5d}x{N:4d}x{K:4d}  f32 {e1:.2e}  f32+= {e2:.2e}  bf16+gelu {e3:.2e}  {'ok' if ok else 'MISMATCH'}", flush=True)
    return bad


def time_shapes():
    side = torch.cuda.Stream()
    print(f"{'M':>6} {'N':>5} {'K':>5} {'epilogue':>8} {'us':>8} {'TFLOP/s':>8}")
    for M in (1024, 2048, 6144, 24576):
        for (N, K) in ((512, 512), (1536, 512), (1024, 512), (512, 1024)):
            for obf, acc, tag in ((1, 0, "bf16"), (0, 1, "f32+=")):
                A = torch.randn(M, K, device=dev).bfloat16()
                Ws = [torch.randn(N, K, device=dev).bfloat16() for _ in range(8)]
                b = torch.randn(N, device=dev)
                out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
                with torch.cuda.stream(side):
                    for W in Ws[:2]:
                        linear(A, W, b, out, acc=acc, st=side.cuda_stream)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        for i in range(24):
                            linear(A, Ws[i % 8], b, out, acc=acc, st=side.cuda_stream)
                    g.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / 240
                print(f"{M:6d} {N:5d} {K:5d} {tag:>8} {us:8.2f} {2.0 * M * N * K / us / 1e6:8.1f}", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        sys.exit(1 if check() else 0)
    time_shapes()
