"""Timeline of one CTA of the attention kernel (bring-up build with -DEDM_ATTN_TRACE, see tools/gpu_attn_trace.sh):
clock64 stamps of the softmax warpgroups and the MMA issuer per kv iteration, printed relative to the iteration's first event."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "libedm_trace.so"))
lib.edm_attention.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.edm_attn_set_trace.argtypes = [C.c_void_p]
B, N, H = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 500, 16
qkv = torch.randn(B * N, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
tr = torch.zeros(96 * 32, device="cuda", dtype=torch.int64)
for it in range(3):
    lib.edm_attn_set_trace(tr.data_ptr() if it == 2 else None)
    assert lib.edm_attention(qkv.data_ptr(), B, N, H, out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
torch.cuda.synchronize()
t = tr.view(96, 32).cpu()
names = ["w0 S seen", "w0 S drained", "w0 O(g-1) seen", "w0 P done", "w1 S seen", "w1 S drained", "w1 O(g-1) seen", "w1 P done",
         "mma S0 issued", "mma S1 issued", "mma PV0 issued", "mma PV1 issued",
         "w0 epi O seen", "w0 epi stored", "w1 epi O seen", "w1 epi stored",
         "mma P0 seen", "mma P1 seen", "mma sfree0 seen", "-", "w0 epi O loaded", "w1 epi O loaded", "w0 P done wp3", "w1 P done wp3"]
base = int(t[t > 0].min())
print("kv-iter " + " ".join(f"{n[-9:]:>9s}" for n in names))
for g in range(20, 28):
    row = t[g]
    print(f"{g:7d} " + " ".join(f"{(int(row[k]) - base) if row[k] > 0 else -1:9d}" for k in range(24)))
