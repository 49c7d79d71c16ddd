import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.parity_utils import full_model
from oracle.weights import make_inputs
cfg, sd, model = full_model()
def run(P, steps, T=60, B=1):
    inp = make_inputs(B, T, P, steps, cfg, seed=77)
    return model.decode_trace(inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"], steps=steps, seed=5)
def cmp(a, b, tag):
    out = []
    for k in ("x0", "x_final", "all_logits", "codes"):
        same = torch.equal(a[k], b[k]) if not a[k].is_floating_point() else bool(((a[k] == b[k]) | (a[k].isnan() & b[k].isnan())).all())
        out.append(f"{k}:{same}")
    for i, (u, v) in enumerate(zip(a["step_logits"], b["step_logits"])):
        out.append(f"step{i}:{torch.equal(u, v)}")
    if not torch.equal(a["all_logits"], b["all_logits"]):
        d = (a["all_logits"] != b["all_logits"]).any(-1)  # [B, Q, T]
        out.append("levels differing: " + str(d.any(-1)[0].tolist()) + " frames: " + str(d.any(1)[0].nonzero().flatten().tolist()[:20]))
    print(tag, " ".join(out))
for (P, steps) in [(0, 1), (20, 1), (0, 3)]:
    a = run(P, steps)
    b = run(P, steps)
    cmp(a, b, f"P={P} S={steps} run1 vs run2:")
    torch.cuda.synchronize()
    model._ws.fill_(0xFF)
    c = run(P, steps)
    cmp(a, c, f"P={P} S={steps} run1 vs poisoned(NaN):")
    model._ws.fill_(0)
    d = run(P, steps)
    cmp(a, d, f"P={P} S={steps} run1 vs zeroed:")
    print("nan in poisoned logits:", c["all_logits"].isnan().any().item())
