#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM through the C ABI (CUDA events, warm, back-to-back launches).
Separates fixed per-launch cost (K=64) from per-K-block cost.  usage (GPU box): python tools/gemm_sweep.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from convofusion_b200 import _lib

dev = "cuda:0"
lib = _lib.lib()


def time_gemm(M, N, K, out_bf16, accumulate, iters=200):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: _lib.check(lib.cfb_linear(A.data_ptr(), 1, W.data_ptr(), b.data_ptr(), out.data_ptr(), int(out_bf16), M, N, K,
                                             0, 0, int(accumulate), _lib.GEMM_TCGEN05, st))
    # eager launches through ctypes cost ~12 us of CPU each, so the launches are captured into a CUDA graph
    # (20 per graph) and the graph is replayed: the events then see device time only.
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        st = side.cuda_stream
        call = lambda: _lib.check(lib.cfb_linear(A.data_ptr(), 1, W.data_ptr(), b.data_ptr(), out.data_ptr(), int(out_bf16),
                                                 M, N, K, 0, 0, int(accumulate), _lib.GEMM_TCGEN05, st))
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(20):
                call()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters // 20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters // 20 * 20)
    return us, 2.0 * M * N * K / us / 1e6


print(f"{'M':>6} {'N':>5} {'K':>5} {'epilogue':>10} {'us':>8} {'TFLOP/s':>8}")
for M in (6144, 1024, 128):
    for N, K in ((512, 64), (512, 512), (512, 1024), (512, 2560), (1536, 512), (2560, 512), (1024, 512), (320, 512), (512, 448)):
        for out_bf16, acc, tag in ((1, 0, "bf16"), (0, 1, "f32+=")):
            if M != 6144 and (N, K) not in ((512, 512), (1536, 512)):
                continue
            us, tf = time_gemm(M, N, K, out_bf16, acc)
            print(f"{M:6d} {N:5d} {K:5d} {tag:>10} {us:8.2f} {tf:8.1f}")
