"""Attention kernel alone: correctness against fp32 SDPA and device time at the bench shape (B=64, N=500, H=16) and a long one.
    python tools/attn_bench.py [path/to/lib.so]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import edm_tts_b200._lib as L  # noqa: E402

if len(sys.argv) > 1:
    L.LIB_PATH = os.path.abspath(sys.argv[1])
lib = L.lib()


def run(B, N, H, iters=50):
    torch.manual_seed(0)
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").to(torch.bfloat16)
    out = torch.zeros(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
    call = lambda: L.check(lib.edm_attention(L.ptr(qkv), B, N, H, L.ptr(out), L.stream_ptr()))
    call()
    q, k, v = qkv[: 2 * N].float().view(2, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(2 * N, H * 64)
    err = (out[: 2 * N].float() - ref).abs().max().item()
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        call()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    # cuDNN / flash SDPA on the same data layout torch prefers ([B, H, N, 64] bf16)
    qb, kb, vb = (t.contiguous() for t in qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4))
    for _ in range(5):
        torch.nn.functional.scaled_dot_product_attention(qb, kb, vb)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        torch.nn.functional.scaled_dot_product_attention(qb, kb, vb)
    b.record()
    torch.cuda.synchronize()
    return ms, a.elapsed_time(b) / iters, err


for B, N, H in ((64, 500, 16), (8, 1650, 16)):
    ms, sdpa, err = run(B, N, H)
    fl = 4.0 * B * H * N * N * 64
    print(f"{os.path.basename(L.LIB_PATH)} B={B} N={N}: {ms * 1e3:.1f} us ({fl / ms / 1e9:.0f} TFLOP/s), torch SDPA {sdpa * 1e3:.1f} us, max |err| {err:.4f}", flush=True)
