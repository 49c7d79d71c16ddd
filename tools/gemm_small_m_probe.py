"""Small-M GEMM time against the number of rows (graph-replayed, so launch gaps do not count): does a row tile with few valid rows
(TMA zero-fills the rest of the 128-row box) cost less?"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import edm_tts_b200._lib as L  # noqa: E402

if os.environ.get("EDM_AB_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["EDM_AB_LIB"])
lib = L.lib()
dev = "cuda"


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for (N, K, epi) in ((4096, 1024, L.EPI_SWISH_BF16), (1024, 4096, L.EPI_RESID_F32)):
    b = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    res = []
    for M in (16, 22, 64, 75, 128, 150, 200, 256):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == L.EPI_RESID_F32 else torch.bfloat16)
        call = lambda: L.check(lib.edm_gemm_bf16(L.ptr(a), K, L.ptr(b), K, M, N, K, epi, L.ptr(bias), L.ptr(out), N, 1e-3, None, None, 1, 0, L.stream_ptr()))
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            call()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                call()
        res.append(f"M={M}: {timeit(g.replay) / 20 * 1e3:.2f}")
    print({k: v for k, v in os.environ.items() if k.startswith("EDM_GEMM")}, f"N={N} K={K}: us per GEMM (20 back-to-back in a graph)  " + "  ".join(res), flush=True)
