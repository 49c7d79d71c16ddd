"""One single-utterance decode (B=1, T=150, 8 steps) for launch-list captures: ncu -k regex:edm:: ... python tools/b1_decode.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edm_tts_b200 import InjectionConformerModel  # noqa: E402
from edm_tts_b200.config import InjectionConformerConfig  # noqa: E402
from edm_tts_b200.synthetic import OracleConfig, make_inputs, make_state_dict  # noqa: E402

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1, 150)
cfg = OracleConfig()
model = InjectionConformerModel(InjectionConformerConfig(), make_state_dict(cfg, 0), device="cuda")
sem = make_inputs(B, T, 0, 1, cfg, seed=1)["semantic_tokens"].cuda()
out = model.infer_special(sem, None, None, steps=8, seed=0)
torch.cuda.synchronize()
print(out.shape, int(out.sum()))
