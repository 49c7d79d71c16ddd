import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.parity_utils import full_model
from oracle.weights import make_inputs
cfg, sd, model = full_model()
def runs(P, steps, n=4, cuda_inputs=True, T=60):
    inp = make_inputs(1, T, P, steps, cfg, seed=77)
    mv = (lambda t: t.cuda() if (t is not None and cuda_inputs) else t)
    sem, ap, sp = mv(inp["semantic_tokens"]), mv(inp["acoustic_prompt_tokens"]), mv(inp["semantic_prompt_tokens"])
    outs = [model.infer_special(sem, ap, sp, steps=steps, seed=5) for _ in range(n)]
    torch.cuda.synchronize()
    tr = model.decode_trace(sem, ap, sp, steps=steps, seed=5)["codes"]
    print(f"P={P} S={steps} cuda_inputs={cuda_inputs}: run_i == run_0:", [torch.equal(o, outs[0]) for o in outs], " == decode_trace:", [torch.equal(o, tr) for o in outs],
          "agree lvl (run1 vs run0):", [round((outs[1][0, q] == outs[0][0, q]).float().mean().item(), 2) for q in range(12)])
runs(20, 3); runs(0, 3); runs(20, 1); runs(0, 1); runs(0, 1, cuda_inputs=False); runs(20, 3, cuda_inputs=False); runs(0, 3)
