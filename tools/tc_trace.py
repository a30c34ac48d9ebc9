#!/usr/bin/env python
"""Phase timeline of one tcgen05 GEMM CTA (library must be built with CFB_EXTRA_NVCC=-DCFB_TC_TRACE=1)."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from convofusion_b200 import _lib
lib = _lib.lib()
lib.cfb_debug_tc_trace.restype, lib.cfb_debug_tc_trace.argtypes = C.c_int, [C.POINTER(C.c_ulonglong)]
names = ["entry", "after init/alloc sync", "first TMA issued", "all TMA issued", "first full barrier", "all MMA issued",
         "acc barrier passed", "TMEM->smem done", "epilogue rows done", "after final sync", "dealloc done"]
dev = "cuda:0"
for (M, N, K, obf, acc) in ((128, 512, 64, 1, 0), (128, 512, 512, 1, 0), (6144, 512, 512, 1, 0), (6144, 512, 512, 0, 1)):
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        _lib.check(lib.cfb_linear(A.data_ptr(), 1, W.data_ptr(), 0, out.data_ptr(), obf, M, N, K, 0, 0, acc, _lib.GEMM_TCGEN05, st))
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    _lib.check(lib.cfb_debug_tc_trace(buf))
    t0 = buf[0]
    print(f"M={M} N={N} K={K} out_bf16={obf} acc={acc}")
    for i, n in enumerate(names):
        print(f"   {n:26s} +{(buf[i] - t0) / 1e3:8.2f} us")
