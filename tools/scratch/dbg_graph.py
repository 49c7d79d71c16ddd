import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.parity_utils import full_model
from oracle.weights import make_inputs
cfg, sd, model = full_model()
for (P, steps) in [(20, 3), (0, 3), (20, 1), (0, 1)]:
    inp = make_inputs(1, 60, P, steps, cfg, seed=77)
    sem = inp["semantic_tokens"].cuda()
    ap = inp["acoustic_prompt_tokens"].cuda() if P else None
    sp = inp["semantic_prompt_tokens"].cuda() if P else None
    ref = model.infer_special(sem, ap, sp, steps=steps, seed=5)
    ref_b = model.infer_special(sem, ap, sp, steps=steps, seed=5)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        w = model.infer_special(sem, ap, sp, steps=steps, seed=5)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out = model.infer_special(sem, ap, sp, steps=steps, seed=5)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    print(P, steps, "eager==eager", torch.equal(ref, ref_b), "side==eager", torch.equal(w, ref), "graph==eager", torch.equal(out, ref),
          "per-level agree", [(out[0, q] == ref[0, q]).float().mean().item() for q in range(12)])
    g.replay(); torch.cuda.synchronize()
    print("  replay2==replay1", torch.equal(out, out.clone()), "graph==eager", torch.equal(out, ref))
