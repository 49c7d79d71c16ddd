"""Functional restatement of the reference S2A decode (test infrastructure; see oracle/__init__.py).

Each function cites the reference lines it follows (paths relative to the reference repo). Two numeric modes:
  mode="fp32"  every op in float32, same op order as the reference on CPU -> pinned tightly to the golden vectors;
  mode="bf16"  the rounding points of the reference under torch.autocast(bfloat16): Linear / Conv /
               matmul / SDPA outputs, Swish, GLU and ChanLayerNorm internals are rounded to bf16, the residual stream,
               LayerNorm and rotary arithmetic stay fp32. This is the arithmetic the CUDA kernels implement.
Weights come as a dict keyed like InjectionConformerModel.state_dict() (oracle/weights.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .weights import OracleConfig, weight_norm_fold


def _r(x: torch.Tensor, mode: str) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if mode == "bf16" else x


def _linear(x, w, b, mode):
    """nn.Linear under autocast: bf16 operands, fp32 accumulate, bf16 result."""
    if mode == "bf16":
        y = _r(x, mode) @ _r(w, mode).t()
        return _r(y + b if b is not None else y, mode)
    return F.linear(x, w, b)


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


# ---------------------------------------------------------------------------------------------- DAC code -> feature
def quantizer_out_weights(sd, cfg: OracleConfig, prefix="acoustic_model.quantizer."):
    """Folded out_proj weights [L, latent, cb_dim], biases [L, latent] and codebooks [L, codes, cb_dim]
    (dac/vector_quantizer.py:28-31, weight-normed 1x1 convs dac/nn_layers.py:8-9)."""
    ws, bs, cbs = [], [], []
    for i in range(cfg.n_codebooks):
        q = f"{prefix}quantizers.{i}."
        w = weight_norm_fold(sd[q + "out_proj.parametrizations.weight.original0"], sd[q + "out_proj.parametrizations.weight.original1"])
        ws.append(w[:, :, 0])
        bs.append(sd[q + "out_proj.bias"])
        cbs.append(sd[q + "codebook.weight"])
    return torch.stack(ws), torch.stack(bs), torch.stack(cbs)


def codes_to_features_unreduced(sd, cfg, codes, mode="fp32"):
    """ResidualVectorQuantize.from_codes_unreduced, dac/vector_quantizer.py:234-252 -> [B, L, latent, T]."""
    w, b, cb = quantizer_out_weights(sd, cfg)
    w, b, cb = w.to(codes.device), b.to(codes.device), cb.to(codes.device)
    out = []
    for i in range(codes.shape[1]):
        z_p = F.embedding(codes[:, i], cb[i])                       # [B, T, cb_dim]  (decode_code :72-73)
        z_q = _linear(z_p, w[i], b[i], mode)                        # out_proj 1x1 conv
        out.append(z_q.transpose(1, 2))
    return torch.stack(out, dim=1)


def codes_to_features(sd, cfg, codes, mode="fp32"):
    """ResidualVectorQuantize.from_codes, dac/vector_quantizer.py:212-232 -> [B, latent, T]."""
    u = codes_to_features_unreduced(sd, cfg, codes, mode)
    z = u[:, 0]
    for i in range(1, u.shape[1]):
        z = _r(z + u[:, i], mode)
    return z


# ---------------------------------------------------------------------------------------------- conformer block
def rotary_freqs(seq_len, dim_head, device):
    """RotaryEmbedding.forward, conformer/conformer.py:28-42."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim_head, 2, device=device).float() / dim_head))
    t = torch.arange(seq_len, device=device).type_as(inv_freq)
    freqs = torch.einsum("i , j -> i j", t, inv_freq)
    return torch.cat((freqs, freqs), dim=-1)


def _rope(pos, t):
    """apply_rotary_pos_emb / rotate_half, conformer/conformer.py:45-51."""
    x1, x2 = t.chunk(2, dim=-1)
    return (t * pos.cos()) + (torch.cat((-x2, x1), dim=-1) * pos.sin())


def feed_forward(sd, p, x, mode):
    """Scale(0.5, PreNorm(FeedForward)), conformer/conformer.py:80-87,102-110,149-157."""
    h = _ln(x, sd[p + "fn.norm.weight"], sd[p + "fn.norm.bias"])
    h = _linear(h, sd[p + "fn.fn.net.0.weight"], sd[p + "fn.fn.net.0.bias"], mode)
    h = _r(h * _r(torch.sigmoid(h), mode), mode)                    # Swish :54-56
    h = _linear(h, sd[p + "fn.fn.net.3.weight"], sd[p + "fn.fn.net.3.bias"], mode)
    return h * 0.5


def attention(sd, p, x, cfg, freqs, mode):
    """PreNorm(Attention), conformer/conformer.py:113-146, Attend.flash_attn attend.py:63-115 (no mask, non-causal)."""
    b, n, _ = x.shape
    h = _ln(x, sd[p + "norm.weight"], sd[p + "norm.bias"])
    q = _linear(h, sd[p + "fn.to_q.weight"], None, mode)
    kv = _linear(h, sd[p + "fn.to_kv.weight"], None, mode)
    k, v = kv.chunk(2, dim=-1)
    q, k, v = (t.view(b, n, cfg.heads, cfg.dim_head).transpose(1, 2) for t in (q, k, v))
    q, k = _r(_rope(freqs, q), mode), _r(_rope(freqs, k), mode)
    o = _r(F.scaled_dot_product_attention(q, k, v), mode)
    o = o.transpose(1, 2).reshape(b, n, cfg.heads * cfg.dim_head)
    return _linear(o, sd[p + "fn.to_out.weight"], sd[p + "fn.to_out.bias"], mode)


def conv_module(sd, p, x, cfg, mode):
    """ConformerConvModule, conformer/conformer.py:160-181 (GLU :59-66, DepthWiseConv1d :69-77 with same padding
    :23-25, Swish, ChanLayerNorm :90-99, pointwise convs as token-major matmuls)."""
    inner = cfg.hidden * cfg.conv_expansion
    h = _ln(x, sd[p + "net.0.weight"], sd[p + "net.0.bias"])
    h = _linear(h, sd[p + "net.2.weight"][:, :, 0], sd[p + "net.2.bias"], mode)       # [B, N, 2*inner]
    out, gate = h.chunk(2, dim=-1)
    h = _r(out * _r(torch.sigmoid(gate), mode), mode)
    pad = cfg.conv_kernel // 2
    pads = (pad, pad - (cfg.conv_kernel + 1) % 2)
    hc = F.pad(h.transpose(1, 2), pads)
    hc = _r(F.conv1d(hc, _r(sd[p + "net.4.conv.weight"], mode), sd[p + "net.4.conv.bias"], groups=inner), mode)   # [B, inner, N]
    hc = _r(hc * _r(torch.sigmoid(hc), mode), mode)
    eps = 1e-4 if mode == "bf16" else 1e-6
    var = _r(torch.var(hc, dim=1, unbiased=False, keepdim=True), mode)
    mean = _r(torch.mean(hc, dim=1, keepdim=True), mode)
    hc = _r(_r(hc - mean, mode) * _r(var.clamp(min=eps).rsqrt(), mode), mode) * sd[p + "net.6.weight"]
    h = hc.transpose(1, 2)
    return _linear(h, sd[p + "net.7.weight"][:, :, 0], sd[p + "net.7.bias"], mode)


def conformer_block(sd, i, x, cfg, freqs, mode, prefix="encoder.layers."):
    """ConformerBlock.forward, conformer/conformer.py:219-235."""
    p = f"{prefix}{i}."
    x = feed_forward(sd, p + "ff1.", x, mode) + x
    x = attention(sd, p + "attn.", x, cfg, freqs, mode) + x
    x = conv_module(sd, p + "conv.", x, cfg, mode) + x
    x = feed_forward(sd, p + "ff2.", x, mode) + x
    return _ln(x, sd[p + "post_norm.weight"], sd[p + "post_norm.bias"])


# ---------------------------------------------------------------------------------------------- injection encoder
def single_to_logits(sd, x, idx, mode):
    """InjectionConformerWrapper.apply_single_to_logits, injection_conformer_wrapper.py:56-63 -> [B, n, codes]."""
    h = _ln(x, sd["encoder.to_logits.0.weight"], sd["encoder.to_logits.0.bias"])
    w = sd["encoder.to_logits.1.weight"][idx]                       # [d, l]
    bias = sd["encoder.to_logits.1.bias"][0, 0, idx]
    if mode == "bf16":
        return _r(_r(_r(h, mode) @ _r(w, mode), mode) + bias, mode)
    return (h @ w) + bias


def forward_first_level(sd, cfg, x, prompt_len=0, mode="fp32"):
    """InjectionConformerWrapper.forward_first_level, injection_conformer_wrapper.py:65-90 -> [B, T, codes] (target rows)."""
    freqs = rotary_freqs(x.shape[-2], cfg.dim_head, x.device)
    for i in range(cfg.depth):
        out = conformer_block(sd, i, x, cfg, freqs, mode)
        if i in cfg.injection_layers:
            return single_to_logits(sd, out, cfg.injection_layers.index(i), mode)[:, prompt_len:]
        x = out
    raise ValueError("no injection layer")


def project_injection(sd, k, inj, mode):
    """project_injection[k] = Linear + LayerNorm, injection_conformer_wrapper.py:26-32."""
    h = _linear(inj, sd[f"encoder.project_injection.{k}.0.weight"], sd[f"encoder.project_injection.{k}.0.bias"], mode)
    return _ln(h, sd[f"encoder.project_injection.{k}.1.weight"], sd[f"encoder.project_injection.{k}.1.bias"])


def forward_full(sd, cfg, x, prompt_injections=None, prompt_len=0, mode="fp32", forced_coarse=None, trace=None):
    """InjectionConformerWrapper.forward in eval mode, injection_conformer_wrapper.py:92-150 -> logits [B, Q, T, codes].
    forced_coarse [B, n_inj, T]: teacher-force the tokens that get injected (parity protocol); logits stay the model's."""
    freqs = rotary_freqs(x.shape[-2], cfg.dim_head, x.device)
    coarse_outs, coarse_logits = [], []
    for i in range(cfg.depth):
        out = conformer_block(sd, i, x, cfg, freqs, mode)
        if i in cfg.injection_layers:
            k = cfg.injection_layers.index(i)
            residual = coarse_outs[-1] if (coarse_outs and cfg.residual) else 0
            coarse_outs.append(out)
            if cfg.use_injection:
                coarse_logits.append(single_to_logits(sd, out, k, mode))            # all rows, [B, N, codes]
                tokens = torch.stack([l.argmax(dim=-1) for l in coarse_logits], dim=1)   # [B, k+1, N]
                if forced_coarse is not None:
                    tokens = tokens.clone()
                    tokens[:, :, prompt_len:] = forced_coarse[:, : k + 1]
                inj = codes_to_features(sd, cfg, tokens, mode).transpose(1, 2)       # [B, N, latent]
                if prompt_injections is not None:
                    is_target = torch.arange(x.shape[1], device=x.device) >= prompt_len
                    inj = torch.where(is_target[None, :, None], inj, prompt_injections[k])
                out = out + project_injection(sd, k, inj, mode) + residual
            else:
                out = out + residual
        x = out
    x_t = x[:, prompt_len:]
    coarse_t = [c[:, prompt_len:] for c in coarse_outs]
    n_fine = cfg.n_codebooks - len(cfg.injection_layers)
    fine = _linear(x_t, sd["encoder.fine_head.0.weight"], sd["encoder.fine_head.0.bias"], mode)
    fine = fine.view(*x_t.shape[:2], n_fine, cfg.hidden)
    allo = torch.cat([c[:, :, None, :] for c in coarse_t] + [fine], dim=-2)            # [B, T, Q, d]
    h = _ln(allo, sd["encoder.to_logits.0.weight"], sd["encoder.to_logits.0.bias"])
    w, bias = sd["encoder.to_logits.1.weight"], sd["encoder.to_logits.1.bias"][0, 0]
    if mode == "bf16":
        logits = _r(_r(torch.einsum("bnqd,qdl->bnql", _r(h, mode), _r(w, mode)), mode) + bias, mode)
    else:
        logits = torch.einsum("bnqd,qdl->bnql", h, w) + bias
    if trace is not None:
        trace["coarse_logits"] = [l[:, prompt_len:] for l in coarse_logits]
    return logits.permute(0, 2, 1, 3)


# ---------------------------------------------------------------------------------------------- decode loop
def random_topk_mask(mask_len, probs, gumbel, temperature, trace=None):
    """edm_tts/utils/utils.py:49-60 with the Gumbel draw injected. trace (dict of lists) receives the confidences and the cut-off."""
    confidence = torch.log(probs) + temperature * gumbel
    sorted_confidence, _ = torch.sort(confidence, dim=-1)
    cut_off = torch.take_along_dim(sorted_confidence, mask_len.long().unsqueeze(-1), dim=-1)
    if trace is not None:
        trace["step_conf"].append(confidence)
        trace["step_cut"].append(cut_off)
    return confidence < cut_off


def build_encoder_input(sd, cfg, semantic_tokens, acoustic_prompt_tokens=None, semantic_prompt_tokens=None, mode="fp32"):
    """InjectionConformerModel.infer_special, modeling_injection_conformer.py:139-168."""
    emb = sd["semantic_embedding.weight"]
    sem = F.embedding(semantic_tokens, emb)
    x = sem + sd["mask_token"].expand(sem.shape[0], sem.shape[1], -1)
    prompt_inj, P = None, 0
    if acoustic_prompt_tokens is not None and semantic_prompt_tokens is not None:
        sem_p = F.embedding(semantic_prompt_tokens, emb)
        u = codes_to_features_unreduced(sd, cfg, acoustic_prompt_tokens, mode)          # [B, q, latent, P]
        ac = _ln(_linear(u[:, 0].transpose(1, 2), sd["acoustic_feat_proj.0.weight"], sd["acoustic_feat_proj.0.bias"], mode),
                 sd["acoustic_feat_proj.1.weight"], sd["acoustic_feat_proj.1.bias"])
        n_inj = min(len(cfg.injection_layers), acoustic_prompt_tokens.shape[1])
        injections = []
        for i in range(n_inj):
            s = u[:, 0]
            for j in range(1, i + 1):
                s = _r(s + u[:, j], mode)
            injections.append(s.transpose(1, 2))
        P = ac.shape[1]
        zeros = torch.zeros(x.shape[0], sem.shape[1], injections[0].shape[-1], device=x.device, dtype=x.dtype)
        prompt_inj = [torch.cat([inj, zeros], dim=1) for inj in injections]
        x = torch.cat([sem_p + ac, x], dim=1)
    return x, sem, prompt_inj, P


def feat_proj(sd, cfg, ids, mode):
    """codes_to_features (1 level) + acoustic_feat_proj, modeling_injection_conformer.py:186-187,193-194."""
    f = codes_to_features(sd, cfg, ids[:, None, :], mode).transpose(1, 2)
    return _ln(_linear(f, sd["acoustic_feat_proj.0.weight"], sd["acoustic_feat_proj.0.bias"], mode),
               sd["acoustic_feat_proj.1.weight"], sd["acoustic_feat_proj.1.bias"])


def infer_special(sd, cfg: OracleConfig, semantic_tokens, acoustic_prompt_tokens=None, semantic_prompt_tokens=None, steps=1,
                  temperature=1.0, cat_gumbel=None, remask_gumbel=None, mode="fp32", forced_ids=None, forced_masks=None,
                  forced_coarse=None, trace=None):
    """InjectionConformerModel.infer_special, modeling_injection_conformer.py:130-230, with the two RNG draws injected:
    cat_gumbel[s] [B*T, codes] (Categorical.sample == argmax(logits + Gumbel)), remask_gumbel[s] [B, T].
    forced_* teacher-force sampled ids [S, B, T], masks after each step [S-1, B, T] and injected coarse tokens.
    trace (dict) receives per-step logits / ids / masks and the final logits."""
    x, sem, prompt_inj, P = build_encoder_input(sd, cfg, semantic_tokens, acoustic_prompt_tokens, semantic_prompt_tokens, mode)
    B, T = semantic_tokens.shape
    mask_token = sd["mask_token"].expand(B, T, -1)
    if trace is not None:
        trace.update(step_logits=[], step_ids=[], step_masks=[], step_conf=[], step_cut=[], x0=x.clone())
    if steps > 1:
        ratios = [math.cos(math.pi / 2.0 * ((t + 1) / steps)) for t in range(steps)]
        mask = torch.ones(B, T, dtype=torch.bool, device=x.device)
        initial = mask.sum(dim=-1)
        for i, ratio in enumerate(ratios):
            logits = forward_first_level(sd, cfg, x.clone(), P, mode)                    # [B, T, codes]
            if i == steps - 1:
                ids = logits.argmax(dim=-1)
            else:
                ids = (logits + cat_gumbel[i].view(B, T, -1).to(logits.device)).argmax(dim=-1)
            if trace is not None:
                trace["step_logits"].append(logits)
                trace["step_ids"].append(ids)
            if forced_ids is not None:
                ids = forced_ids[i]
            feats = feat_proj(sd, cfg, ids, mode)
            x[:, P:] = torch.where(mask[..., None], sem + feats, x[:, P:])
            if i < steps - 1:
                mask_len = torch.floor(initial * ratio)
                mask_len = torch.maximum(torch.ones_like(mask_len), torch.minimum(torch.sum(mask, dim=-1) - 1, mask_len))
                probs = F.softmax(logits, dim=-1)
                sel = torch.take_along_dim(probs, ids.unsqueeze(-1), -1).squeeze(-1)
                sel = torch.where(mask, sel, torch.inf)
                next_mask = random_topk_mask(mask_len, sel, remask_gumbel[i].to(sel.device), temperature * ratio, trace)
                if trace is not None:
                    trace["step_masks"].append(next_mask)
                if forced_masks is not None:
                    next_mask = forced_masks[i]
                x[:, P:] = torch.where(next_mask[..., None], sem + mask_token, x[:, P:])
                mask = next_mask
    if trace is not None:
        trace["x_final"] = x.clone()
    all_logits = forward_full(sd, cfg, x, prompt_inj, P, mode, forced_coarse, trace)
    if trace is not None:
        trace["all_logits"] = all_logits
    return all_logits.argmax(dim=-1)


def training_forward(sd, cfg: OracleConfig, acoustic_tokens, semantic_tokens, mask_time_indices, mode="fp32", loss_all=False):
    """InjectionConformerModel.forward in eval mode, modeling_injection_conformer.py:76-128, with the cosine_schedule_mask draw
    (:62-74) injected. Eval mode of the wrapper (injection_conformer_wrapper.py:113-131): the coarse logits are predicted, but
    the features injected at every row are the ground-truth ones (`injections[injection_idx]`).
    Returns loss, output codes (flat over the masked (b, q, t) positions, or [b, q, t] with loss_all, as the reference returns them) and all logits."""
    assert acoustic_tokens.shape[-1] == semantic_tokens.shape[-1], "Acoustic and semantic tokens must have same length"
    sem = F.embedding(semantic_tokens, sd["semantic_embedding.weight"])
    B, T, _ = sem.shape
    ac = feat_proj(sd, cfg, acoustic_tokens[:, 0], mode)                                     # :92-94
    m = mask_time_indices.bool()
    x = torch.where(m[:, :, None], sem + sd["mask_token"].expand(B, T, -1), sem + ac)        # :100-102
    logits = forward_full(sd, cfg, x, mode=mode, forced_coarse=acoustic_tokens[:, : len(cfg.injection_layers)])   # [B, Q, T, V]
    V = logits.shape[-1]
    if not loss_all:
        sel = logits.masked_select(m[:, None, :, None]).view(-1, V)                          # :114-116
        tgt = acoustic_tokens.masked_select(m[:, None, :]).view(-1)
    else:
        sel, tgt = logits, acoustic_tokens.reshape(-1)                                       # :117-118: logits keep [b, q, t, V]
    loss = F.cross_entropy(sel.reshape(-1, V).float(), tgt, reduction="mean")                # :120-122
    return dict(loss=loss, output_acoustic_codes=sel.argmax(dim=-1), target_acoustic_codes=acoustic_tokens, logits=logits)
