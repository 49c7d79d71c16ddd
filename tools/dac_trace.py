"""Timeline of one CTA of the fused 128-channel ResidualUnit kernel (bring-up build with -DEDM_DAC_TRACE, tools/gpu_dac_trace.sh)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = C.CDLL(os.path.join(ROOT, "gpurun_out", "libedm_trace.so"))
c, B, rows = 128, 2, 480080
dev = "cuda"
a = torch.randn(B, rows, c, device=dev).to(torch.bfloat16)
so = torch.empty_like(a)
y = torch.randn(B, rows, c, device=dev)
w7 = (torch.randn(c, 7 * c, device=dev) * 0.03).to(torch.bfloat16)
w1 = (torch.randn(c, c, device=dev) * 0.05).to(torch.bfloat16)
v = [torch.rand(c, device=dev) + 0.5 for _ in range(4)]
tr = torch.zeros(64 * 16, device=dev, dtype=torch.int64)
lib.edm_dac_set_trace.argtypes = [C.c_void_p]
lib.edm_dac_resunit.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
for it in range(3):
    lib.edm_dac_set_trace(tr.data_ptr() if it == 2 else None)
    rc = lib.edm_dac_resunit(a.data_ptr(), rows * c, B, rows, c, 3, w7.data_ptr(), w1.data_ptr(), v[0].data_ptr(), v[1].data_ptr(), v[2].data_ptr(), v[3].data_ptr(),
                             y.data_ptr(), rows * c, so.data_ptr(), rows * c, 0, rows, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
t = tr.view(64, 16).cpu()
names = ["g1 start", "g1 issued", "hfull seen", "t2empty ok", "g2 issued", "p1 hfree", "p1 t1full", "p1 done", "p2 wait", "p2 t2full", "c0 tmem", "c0 x in", "c0 stored",
         "c0 y sts", "c0 s sts", "c0 fenced"]
base = int(t[t > 0].min())
print("tile " + " ".join(f"{n:>10s}" for n in names))
for tl in range(20, 30):
    print(f"{tl:4d} " + " ".join(f"{(int(t[tl][k]) - base) if t[tl][k] > 0 else -1:10d}" for k in range(16)))
