"""Where the time of a single-utterance decode goes (BASELINE config 1 shape on the GPU: B=1, T=150, 8 steps): device time of the
whole call, host time of the launch loop, and per-kernel-class device times from the library's event instrumentation."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edm_tts_b200 import InjectionConformerModel, _lib  # noqa: E402
from edm_tts_b200.config import InjectionConformerConfig  # noqa: E402
from edm_tts_b200.synthetic import OracleConfig, make_inputs, make_state_dict  # noqa: E402

cfg = OracleConfig()
model = InjectionConformerModel(InjectionConformerConfig(), make_state_dict(cfg, 0), device="cuda")
lib = _lib.lib()
for B, T in [(1, 150), (1, 500), (4, 500)]:
    sem = make_inputs(B, T, 0, 1, cfg, seed=1)["semantic_tokens"].cuda()
    for _ in range(2):
        model.infer_special(sem, None, None, steps=8, seed=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    model.infer_special(sem, None, None, steps=8, seed=0)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n0 = lib.edm_launch_count()
    model.infer_special(sem, None, None, steps=8, seed=0)
    torch.cuda.synchronize()
    launches = lib.edm_launch_count() - n0
    lib.edm_prof_enable(1)
    model.infer_special(sem, None, None, steps=8, seed=0)
    torch.cuda.synchronize()
    pm, pw, pc = (C.c_double * 8)(), (C.c_double * 8)(), (C.c_int * 8)()
    lib.edm_prof_collect(pm, pw, pc)
    lib.edm_prof_enable(0)
    names = ["gemm", "attention", "layernorm", "conv_module"]
    per = ", ".join(f"{n} {pc[i]} x {pm[i] / max(pc[i], 1) * 1e3:.1f} us = {pm[i]:.2f} ms" for i, n in enumerate(names))
    print(f"B={B} T={T}: device {e0.elapsed_time(e1):.2f} ms, host launch loop {(t1 - t0) * 1e3:.2f} ms, wall {(t2 - t0) * 1e3:.2f} ms, {launches} launches; {per}", flush=True)

# What a captured graph of the same call would give (fixed seed: measurement only)
for B, T in [(1, 150), (1, 500)]:
    sem = make_inputs(B, T, 0, 1, cfg, seed=1)["semantic_tokens"].cuda()
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(2):
                model.infer_special(sem, None, None, steps=8, seed=0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = model.infer_special(sem, None, None, steps=8, seed=0)
        torch.cuda.synchronize()
        ref = model.infer_special(sem, None, None, steps=8, seed=0)
        g.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(out, ref))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"B={B} T={T}: graph replay {e0.elapsed_time(e1) / 5:.2f} ms per decode (codes equal to the stream launch: {same})", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"B={B} T={T}: graph capture failed: {type(ex).__name__}: {str(ex)[:200]}", flush=True)
