"""Sustained decode throughput as a function of the per-launch batch (chunk) size at T = 500, S = 8: smaller chunks keep more of
the inter-kernel traffic in the 126 MB L2 (less DRAM energy under the power cap) but quantise the GEMM tile waves worse.
Each size runs back to back for ~4 s; frames/s over the second half."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from edm_tts_b200 import InjectionConformerModel  # noqa: E402
from edm_tts_b200.config import InjectionConformerConfig  # noqa: E402
from edm_tts_b200.synthetic import OracleConfig, make_inputs, make_state_dict  # noqa: E402

cfg = OracleConfig()
model = InjectionConformerModel(InjectionConformerConfig(), make_state_dict(cfg, 0), device="cuda")
sizes = [int(a) for a in sys.argv[1:]] or [64, 37, 32, 24, 16]
for B in sizes:
    T = 500
    sem = make_inputs(B, T, 0, 1, cfg, seed=1)["semantic_tokens"].cuda()
    for _ in range(2):
        model.infer_special(sem, None, None, steps=8, seed=0)
    torch.cuda.synchronize()
    t_end = time.time() + 4.0
    n, n_mid = 0, None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    while time.time() < t_end:
        if n_mid is None and time.time() > t_end - 2.0:
            e0.record()
            n_mid = n
        model.infer_special(sem, None, None, steps=8, seed=0)
        n += 1
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / max(1, n - n_mid)
    print(json.dumps({"B": B, "T": T, "ms": ms, "frames_per_s": B * T / (ms * 1e-3)}), flush=True)
