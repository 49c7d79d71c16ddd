"""Text-to-semantic decode alone (the bench's secondary_t2s shape): ms per utterance.  EDM_AB_LIB=path/to/other.so for another build."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import edm_tts_b200._lib as L  # noqa: E402

if os.environ.get("EDM_AB_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["EDM_AB_LIB"])
    if os.environ.get("EDM_AB_OLD"):
        import ctypes as _C

        _h = _C.CDLL(L.LIB_PATH)
        for _n in list(L._SIGNATURES):
            if not hasattr(_h, _n):
                del L._SIGNATURES[_n]
from edm_tts_b200 import TextToSemanticWLen  # noqa: E402
from edm_tts_b200.config import TextToSemanticWLenConfig  # noqa: E402
from edm_tts_b200.synthetic import T2SConfig, make_t2s_state_dict  # noqa: E402

dims = T2SConfig(hidden=384, heads=8, depth=12, lp_heads=8, lp_depth=4)
t2s = TextToSemanticWLen(TextToSemanticWLenConfig(hidden_size=384, main_encoder_args=dict(depth=12, heads=8), length_predictor_args=dict(depth=4, heads=8)),
                         make_t2s_state_dict(dims, 0), device="cuda", max_positions=1024)
text = "The quick brown fox jumps over the lazy dog, and the dog, for once, does not mind at all."
T = 500
for _ in range(3):
    t2s.infer(text, pred_iters=16, gt_length=T, seed=1)
torch.cuda.synchronize()
res = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        t2s.infer(text, pred_iters=16, gt_length=T, seed=1)
    b.record()
    torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 5)
print(os.path.basename(L.LIB_PATH), "text-to-semantic ms per utterance:", " ".join(f"{r:.3f}" for r in res), flush=True)
