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


def time_concurrent(M, N, K, out_bf16, accumulate, n_streams, reps=10, iters=10):
    """n_streams independent chains of `reps` dependent GEMMs each (distinct buffers per chain, shared weights),
    forked/joined inside one CUDA graph: the aggregate throughput the sampler's concurrent chains can reach."""
    W = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev)
    As = [torch.randn(M, K, device=dev).bfloat16() for _ in range(n_streams)]
    outs = [torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32) for _ in range(n_streams)]
    main = torch.cuda.Stream()
    subs = [torch.cuda.Stream() for _ in range(n_streams)]

    def body():
        ev = torch.cuda.Event()
        ev.record(main)
        for s, A, o in zip(subs, As, outs):
            s.wait_event(ev)
            for _ in range(reps):
                _lib.check(lib.cfb_linear(A.data_ptr(), 1, W.data_ptr(), b.data_ptr(), o.data_ptr(), int(out_bf16), M, N, K,
                                          0, 0, int(accumulate), _lib.GEMM_TCGEN05, s.cuda_stream))
            e = torch.cuda.Event()
            e.record(s)
            main.wait_event(e)

    with torch.cuda.stream(main):
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            body()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    n = n_streams * reps
    return us / reps, 2.0 * M * N * K * n / us / 1e6


if "--concurrent" in sys.argv:
    print(f"\n{'streams':>7} {'M':>6} {'N':>5} {'K':>5} {'epilogue':>10} {'us/round':>9} {'agg TFLOP/s':>12}")
    for S in (1, 2, 4, 6, 12):
        for (M, N, K) in ((1056, 512, 512), (1056, 1536, 512), (1056, 512, 1024)):
            for out_bf16, acc, tag in ((1, 0, "bf16"), (0, 1, "f32+=")):
                us, tf = time_concurrent(M, N, K, out_bf16, acc, S)
                print(f"{S:7d} {M:6d} {N:5d} {K:5d} {tag:>10} {us:9.2f} {tf:12.1f}")
