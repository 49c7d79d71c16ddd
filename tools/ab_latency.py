"""Single-utterance decode latency, default mode against low-latency mode (set_low_latency), on the product library."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import edm_tts_b200._lib as L  # noqa: E402

if os.environ.get("EDM_AB_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["EDM_AB_LIB"])
from edm_tts_b200 import InjectionConformerModel  # noqa: E402
from edm_tts_b200.config import InjectionConformerConfig  # noqa: E402
from edm_tts_b200.synthetic import OracleConfig, make_state_dict  # noqa: E402

model = InjectionConformerModel(InjectionConformerConfig(), make_state_dict(OracleConfig(), 0))
sem = torch.randint(0, 1024, (4, 500), device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for mode in (False, True):
    model.set_low_latency(mode)
    res = []
    for B, T in ((1, 150), (1, 256), (1, 500), (2, 150), (2, 500), (4, 500)):
        tok = sem[:B, :T].contiguous()
        res.append(f"B={B},T={T}: {timeit(lambda: model.infer_special(tok, None, None, steps=8, seed=1)):.3f}")
    print({k: v for k, v in os.environ.items() if k.startswith("EDM_SPLIT")}, f"low_latency={mode} ms per decode:", "  ".join(res), flush=True)
