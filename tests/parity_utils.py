"""Shared helpers of the GPU parity tests: build identical weights for the oracle and the CUDA path, run both under the
teacher-forcing protocol and classify arg-max disagreements as near-ties by the oracle's margin."""
from __future__ import annotations

import torch

from oracle import s2a as os2a
from oracle.weights import OracleConfig, make_inputs, make_state_dict

NEAR_TIE_EPS = 0.12  # a disagreement is a documented near-tie iff oracle_logit[its choice] - oracle_logit[our choice] < eps

_CACHE = {}


def full_model(seed=0, device="cuda"):
    """(OracleConfig, oracle state dict on `device`, CUDA-path model) for the reference's base config."""
    key = (seed, device)
    if key not in _CACHE:
        from edm_tts_b200 import InjectionConformerModel
        from edm_tts_b200.config import InjectionConformerConfig

        cfg = OracleConfig()
        sd_cpu = make_state_dict(cfg, seed)
        model = InjectionConformerModel(InjectionConformerConfig(), sd_cpu, device=device)
        sd = {k: v.to(device) for k, v in sd_cpu.items()}
        _CACHE[key] = (cfg, sd, model)
    return _CACHE[key]


def oracle_trace(cfg, sd, inp, steps, temperature=1.0, mode="fp32", device="cuda", **forced):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    to = lambda t: None if t is None else t.to(device)
    trace = {}
    with torch.inference_mode():
        codes = os2a.infer_special(sd, cfg, to(inp["semantic_tokens"]), to(inp["acoustic_prompt_tokens"]), to(inp["semantic_prompt_tokens"]),
                                   steps=steps, temperature=temperature, cat_gumbel=to(inp["cat_gumbel"]),
                                   remask_gumbel=to(inp["remask_gumbel"]), mode=mode, trace=trace, **forced)
    trace["codes"] = codes
    return trace


def compare_logits(ours: torch.Tensor, ref: torch.Tensor, what: str, eps=NEAR_TIE_EPS):
    """-> dict(max, mean, agree, n_mismatch, n_not_near_tie). ours/ref [..., V]."""
    d = (ours.float() - ref.float()).abs()
    a_o, a_r = ours.argmax(-1), ref.argmax(-1)
    mism = a_o != a_r
    margin = ref.gather(-1, a_r[..., None])[..., 0] - ref.gather(-1, a_o[..., None])[..., 0]
    bad = mism & (margin >= eps)
    return dict(what=what, max=d.max().item(), mean=d.mean().item(), agree=1.0 - mism.float().mean().item(), n=mism.numel(),
                n_mismatch=int(mism.sum().item()), n_not_near_tie=int(bad.sum().item()),
                worst_margin=margin[mism].max().item() if mism.any() else 0.0)


def teacher_forced_parity(cfg, sd, model, B, T, P, steps, input_seed, temperature=1.0, mode="fp32"):
    """Runs the oracle free, then the CUDA path with the oracle's decisions forced upstream of every graded stage."""
    inp = make_inputs(B, T, P, steps, cfg, seed=input_seed)
    ref = oracle_trace(cfg, sd, inp, steps, temperature, mode)
    n_inj = len(cfg.injection_layers)
    forced = dict(forced_coarse=ref["all_logits"][:, :n_inj].argmax(-1))
    if steps > 1:
        forced["forced_ids"] = torch.stack(ref["step_ids"])
        forced["forced_masks"] = torch.stack(ref["step_masks"])
    ours = model.decode_trace(inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"], steps=steps,
                              temperature=temperature, cat_gumbel=inp["cat_gumbel"] if steps > 1 else None,
                              remask_gumbel=inp["remask_gumbel"] if steps > 1 else None, **forced)
    torch.cuda.synchronize()
    reports = []
    for s in range(len(ref["step_logits"])):
        reports.append(compare_logits(ours["step_logits"][s], ref["step_logits"][s], f"first-level logits step {s}"))
    for q in range(cfg.n_codebooks):
        reports.append(compare_logits(ours["all_logits"][:, q], ref["all_logits"][:, q], f"final logits level {q}"))
    return inp, ref, ours, reports
