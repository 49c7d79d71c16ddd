"""Functional restatement of the reference text-to-semantic decode, TextToSemanticWLen.infer
(edm_tts/models/text_to_semantic/modeling_text_to_semantic.py:184-267), with its two RNG draws injected. Test infrastructure
(see oracle/__init__.py); pinned to the unmodified reference by tests/golden/t2s_*.pt (tests/test_oracle_golden.py).

The conformer blocks are the ones of oracle/s2a.py (same ConformerBlock class in the reference, conformer/conformer.py:184-291).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import s2a as os2a

PAD, TEXT, SPEECH, SEP, MASK = 0, 1, 2, 3, 4      # configuration.py:45-51


def _stack_cfg(cfg, heads):
    return SimpleNamespace(hidden=cfg.hidden, heads=heads, dim_head=cfg.hidden // heads, conv_expansion=cfg.conv_expansion,
                           conv_kernel=cfg.conv_kernel)


def conformer(sd, prefix, depth, heads, x, cfg, mode="fp32"):
    """Conformer.forward without mask, conformer/conformer.py:271-291: rotary embedding + `depth` blocks."""
    c = _stack_cfg(cfg, heads)
    freqs = os2a.rotary_freqs(x.shape[-2], c.dim_head, x.device)
    for i in range(depth):
        x = os2a.conformer_block(sd, i, x, c, freqs, mode, prefix=prefix + "layers.")
    return x


def text_tokens_of(text: str, cfg, device="cpu"):
    """:193-194: utf-8 bytes shifted past the special tokens."""
    return torch.tensor(list(text.encode("utf-8")), dtype=torch.long, device=device) + cfg.num_special


def predict_length(sd, cfg, text_tokens, mode="fp32", return_raw=False):
    """:198-203: length predictor on [length_token, text embeddings], position 0 -> Linear(d, 1) -> exp -> ceil."""
    emb = F.embedding(text_tokens, sd["input_embedding.weight"]).unsqueeze(0)
    x = torch.cat([sd["length_token"], emb], dim=1)
    out = conformer(sd, "length_predictor.", cfg.lp_depth, cfg.lp_heads, x, cfg, mode)[:, 0]
    raw = os2a._linear(out, sd["length_pred_head.weight"], sd["length_pred_head.bias"], mode).squeeze(-1)      # log-length
    length = raw.exp().ceil().long()
    return (length, raw) if return_raw else length


def embeddings_to_logits(sd, cfg, x, mode="fp32"):
    """:135-152 with mask=None: main conformer -> pred_transform (Linear, GELU(tanh), LayerNorm) -> pred_head. -> [1, L, V]"""
    out = conformer(sd, "conformer.", cfg.depth, cfg.heads, x, cfg, mode)
    h = os2a._linear(out, sd["pred_transform.0.weight"], sd["pred_transform.0.bias"], mode)
    h = os2a._r(F.gelu(h, approximate="tanh"), mode)
    h = F.layer_norm(h, (h.shape[-1],), sd["pred_transform.2.weight"], sd["pred_transform.2.bias"], 1e-5)
    return os2a._linear(h, sd["pred_head.weight"], sd["pred_head.bias"], mode)


def build_sequence(cfg, text_tokens, length):
    """:205-217: [text] bytes [sep] [speech] [mask]*length [sep] and the mask of the speech positions."""
    dev = text_tokens.device
    one = lambda v: torch.tensor([v], dtype=torch.long, device=dev)
    input_ids = torch.cat([one(TEXT), text_tokens, one(SEP), one(SPEECH), one(MASK).repeat(int(length)), one(SEP)]).unsqueeze(0)
    full_mask = torch.zeros_like(input_ids, dtype=torch.bool)
    start = text_tokens.numel() + 3
    full_mask[0, start:start + int(length)] = True
    return input_ids, full_mask


def infer(sd, cfg, text, pred_iters=10, temperature=1.0, gt_length=None, cat_gumbel=None, remask_gumbel=None, mode="fp32",
          forced_ids=None, forced_masks=None, trace=None, device="cpu"):
    """TextToSemanticWLen.infer, :184-267 -> speech_pred_tokens [length] (semantic vocabulary).
    cat_gumbel[i] [L, V]: Categorical(logits).sample() == argmax(logits + g); remask_gumbel[i] [1, L]: the Gumbel draw of
    random_topk_mask. forced_ids [iters, 1, L] / forced_masks [iters-1, 1, L] teacher-force the sampled tokens / next masks (parity
    protocol); trace receives per-iteration logits, own ids, own masks, confidences and cut-offs."""
    text_tokens = text if torch.is_tensor(text) else text_tokens_of(text, cfg, device)
    if gt_length is not None:
        length = int(gt_length)
    else:
        length = int(predict_length(sd, cfg, text_tokens, mode).item())
    input_ids, full_mask = build_sequence(cfg, text_tokens, length)
    offset = cfg.num_special + cfg.text_vocab
    sampled = input_ids.clone()
    if trace is not None:
        trace.update(step_logits=[], step_ids=[], step_masks=[], step_conf=[], step_cut=[], input_ids=input_ids, full_mask=full_mask, length=length)
    ratios = [math.cos(math.pi / 2.0 * ((t + 1) / pred_iters)) for t in range(pred_iters)]
    mask = full_mask.clone()
    initial = mask.sum(dim=-1)
    emb = sd["input_embedding.weight"]
    for i, ratio in enumerate(ratios):
        logits = embeddings_to_logits(sd, cfg, F.embedding(sampled, emb), mode)                      # [1, L, V]
        if i == pred_iters - 1:
            ids = logits.argmax(dim=-1)
            if trace is not None:
                trace["step_logits"].append(logits)
                trace["step_ids"].append(ids)
            if forced_ids is not None:
                ids = forced_ids[i]
            sampled = torch.where(full_mask, ids, input_ids)                                          # :231-233
        else:
            ids = (logits + cat_gumbel[i].view(1, -1, logits.shape[-1]).to(logits.device)).argmax(dim=-1)   # :235
            if trace is not None:
                trace["step_logits"].append(logits)
                trace["step_ids"].append(ids)
            if forced_ids is not None:
                ids = forced_ids[i]
            mask_len = torch.floor(initial.float() * ratio).long()                                    # :237
            mask_len = torch.maximum(torch.tensor(1, device=logits.device), torch.minimum(mask_len, initial))   # :240-242
            probs = F.softmax(logits, dim=-1)
            sel = torch.take_along_dim(probs, ids.unsqueeze(-1), -1).squeeze(-1)
            sel = torch.where(mask, sel, torch.inf)                                                   # :250
            next_mask = os2a.random_topk_mask(mask_len, sel, remask_gumbel[i].to(sel.device), temperature * ratio, trace)
            if trace is not None:
                trace["step_masks"].append(next_mask)
            if forced_masks is not None:
                next_mask = forced_masks[i]
            sampled = torch.where(next_mask, torch.tensor(MASK, device=ids.device), ids + offset)     # :256-257
            sampled = torch.where(full_mask, sampled, input_ids)                                      # :258
            mask = next_mask
    return sampled[full_mask]
