#!/usr/bin/env python
"""GEMM + LayerNorm: separate row kernel vs the GEMM's LayerNorm tail (debug entry), graph-captured chains.
With a library built with -DCFB_TC_TRACE=1 also prints the phase timeline of CTA (0,0)."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from convofusion_b200 import _lib
lib = _lib.lib()
P, I = C.c_void_p, C.c_int
lib.cfb_debug_linear_ln_tail.restype = C.c_int
lib.cfb_debug_linear_ln_tail.argtypes = [P, P, P, P, P, P, P, P, I, I, P]
lib.cfb_debug_tc_trace.restype, lib.cfb_debug_tc_trace.argtypes = C.c_int, [C.POINTER(C.c_ulonglong)]
dev = "cuda:0"


def bench(M, K, n_streams, fused, reps=10, iters=10):
    W = (torch.randn(512, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(512, device=dev)
    g, b = torch.rand(512, device=dev) + 0.5, torch.randn(512, device=dev)
    As = [torch.randn(M, K, device=dev).bfloat16() for _ in range(n_streams)]
    hs = [torch.randn(M, 512, device=dev) for _ in range(n_streams)]
    outs = [torch.zeros(M, 512, device=dev, dtype=torch.bfloat16) for _ in range(n_streams)]
    cnts = [torch.zeros(2 * (M // 128 + 2), device=dev, dtype=torch.int32) for _ in range(n_streams)]
    main = torch.cuda.Stream()
    subs = [torch.cuda.Stream() for _ in range(n_streams)]

    def one(A, h, o, c, st):
        if fused:
            _lib.check(lib.cfb_debug_linear_ln_tail(A.data_ptr(), W.data_ptr(), bias.data_ptr(), h.data_ptr(), o.data_ptr(),
                                                    g.data_ptr(), b.data_ptr(), c.data_ptr(), M, K, st))
        else:
            _lib.check(lib.cfb_linear(A.data_ptr(), 1, W.data_ptr(), bias.data_ptr(), h.data_ptr(), 0, M, 512, K, 0, 0, 1,
                                      _lib.GEMM_TCGEN05, st))
            _lib.check(lib.cfb_layernorm(h.data_ptr(), g.data_ptr(), b.data_ptr(), o.data_ptr(), 1, M, 512, st))

    def body():
        ev = torch.cuda.Event(); ev.record(main)
        for s, A, h, o, c in zip(subs, As, hs, outs, cnts):
            s.wait_event(ev)
            for _ in range(reps):
                one(A, h, o, c, s.cuda_stream)
            e = torch.cuda.Event(); e.record(s); main.wait_event(e)

    with torch.cuda.stream(main):
        body(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=main):
            body()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters / reps, (hs[0], outs[0], g, b)


# correctness of the tail against torch
us, (h, o, g, b) = bench(1024, 512, 1, True, reps=1, iters=1)
ref = torch.nn.functional.layer_norm(h, (512,), g, b)
print("tail vs torch layer_norm: max abs diff", float((o.float() - ref).abs().max()))
names = {0: "entry", 1: "after init", 5: "all MMA issued", 6: "acc ready", 7: "tile handed to TMA", 9: "after sync",
         10: "dealloc", 11: "arrived on counter", 12: "block complete", 13: "rows normalised"}
buf = (C.c_ulonglong * 16)()
_lib.check(lib.cfb_debug_tc_trace(buf))
if buf[0]:
    for i in sorted(names):
        print(f"   {names[i]:22s} +{(buf[i] - buf[0]) / 1e3:8.2f} us")
print(f"{'streams':>7} {'M':>6} {'K':>5} {'separate us':>12} {'tail us':>9}")
for S in (1, 6, 12):
    for M, K in ((1024, 512), (1024, 1024)):
        a, _ = bench(M, K, S, False)
        t, _ = bench(M, K, S, True)
        print(f"{S:7d} {M:6d} {K:5d} {a:12.2f} {t:9.2f}")
