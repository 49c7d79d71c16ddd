"""A/B timing of the bench step (B=64 x 500 frames x 8 steps) and of the single-utterance decode under the bring-up library
(gpurun_out/libedm_bringup.so, built with -DEDM_BRINGUP so that the EDM_* environment switches exist).
    EDM_PDL=0 python tools/ab_step.py [steps]          EDM_AB_LIB=path/to/other.so python tools/ab_step.py   (another build of the library)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import edm_tts_b200._lib as L  # noqa: E402

alt = os.environ.get("EDM_AB_LIB") or os.path.join(ROOT, "gpurun_out", "libedm_bringup.so")
if os.path.exists(alt):
    L.LIB_PATH = os.path.abspath(alt)
from edm_tts_b200 import InjectionConformerModel  # noqa: E402
from edm_tts_b200.config import InjectionConformerConfig  # noqa: E402
from edm_tts_b200.synthetic import OracleConfig, make_state_dict  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cfg = OracleConfig()
model = InjectionConformerModel(InjectionConformerConfig(), make_state_dict(cfg, 0))
sem = torch.randint(0, 1024, (64, 500), device="cuda")


def timeit(fn, n):
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


res = []
for B, T, n in ((1, 150, 20), (1, 500, 20), (4, 500, 10), (8, 500, 10), (16, 500, 5), (32, 500, 5), (64, 500, steps)):
    tok = sem[:B, :T].contiguous()
    res.append(f"B={B},T={T}: {timeit(lambda: model.infer_special(tok, None, None, steps=8, seed=1), n):.3f}")
print({k: v for k, v in os.environ.items() if k.startswith("EDM_")}, "ms per decode:", "  ".join(res), flush=True)
