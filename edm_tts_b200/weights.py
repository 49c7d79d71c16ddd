"""Packs a reference state dict (InjectionConformerModel.state_dict()) into the tensors the C ABI
names (edm_s2a_weight_name): bf16 K-major GEMM operands, fp32 norms / biases, and the folded lookup tables.

Folding (exact algebra, evaluated in float64 then stored as fp32):
  proj[i][code]        = W_out_i codebook_i[code]                      (weight-norm folded, dac/vector_quantizer.py:28-31)
  feat_table[code]     = W_fp proj[0][code],  feat_const = W_fp b_out_0 + b_fp      (acoustic_feat_proj.0 o from_codes)
  inj_table[k][i][code]= W_pi_k proj[i][code], inj_const[k] = W_pi_k sum_{i<=k} b_out_i + b_pi_k   (project_injection[k].0)
so "codes -> DAC feature -> Linear" becomes a gather-sum; the LayerNorm that follows runs in the kernel.
"""
from __future__ import annotations

import torch

from .config import InjectionConformerConfig


def weight_norm_fold(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """w = g * v / ||v|| (norm over all dims but 0) -- torch.nn.utils.parametrizations.weight_norm, dac/nn_layers.py:8-9."""
    return v * (g / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1))))


def quantizer_tensors(sd: dict, n_codebooks: int, prefix: str, device, dtype=torch.float64):
    """Folded in/out projections and codebooks of the DAC quantizer as dense tensors."""
    w_in, b_in, w_out, b_out, cb = [], [], [], [], []
    for i in range(n_codebooks):
        q = f"{prefix}quantizers.{i}."
        g_in, v_in = sd[q + "in_proj.parametrizations.weight.original0"], sd[q + "in_proj.parametrizations.weight.original1"]
        g_out, v_out = sd[q + "out_proj.parametrizations.weight.original0"], sd[q + "out_proj.parametrizations.weight.original1"]
        w_in.append(weight_norm_fold(g_in.to(device, dtype), v_in.to(device, dtype))[:, :, 0])      # [cb_dim, latent]
        w_out.append(weight_norm_fold(g_out.to(device, dtype), v_out.to(device, dtype))[:, :, 0])   # [latent, cb_dim]
        b_in.append(sd[q + "in_proj.bias"].to(device, dtype))
        b_out.append(sd[q + "out_proj.bias"].to(device, dtype))
        cb.append(sd[q + "codebook.weight"].to(device, dtype))
    return torch.stack(w_in), torch.stack(b_in), torch.stack(w_out), torch.stack(b_out), torch.stack(cb)


def glu_interleave(n_out: int) -> torch.Tensor:
    """Row order [v0..v31, g0..g31, v32..v63, g32..g63, ...] for a [2C] -> (value C | gate C) projection."""
    c = n_out // 2
    v = torch.arange(c).view(-1, 32)
    return torch.cat([v, v + c], dim=1).reshape(-1)


def rope_tables(max_positions: int, dim_head: int, device):
    """cos / sin of RotaryEmbedding.forward (conformer/conformer.py:28-42); only the first half is stored (freqs = cat(f, f))."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim_head, 2, device=device).float() / dim_head))
    t = torch.arange(max_positions, device=device).type_as(inv_freq)
    freqs = torch.einsum("i , j -> i j", t, inv_freq)
    return freqs.cos().contiguous(), freqs.sin().contiguous()


def pack_s2a_weights(sd: dict, cfg: InjectionConformerConfig, device, max_positions: int = 4096) -> dict:
    """-> {abi_name: contiguous CUDA tensor}. Raises KeyError if the state dict lacks a hot-path key (strict, like the
    reference's load_state_dict(strict=True))."""
    d = cfg.hidden_size
    f32 = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
    bf16 = lambda t: t.to(device=device, dtype=torch.float32).to(torch.bfloat16).contiguous()
    out = {}
    for i in range(cfg.depth):
        p, o = f"encoder.layers.{i}.", f"blocks.{i}."
        for ff in ("ff1", "ff2"):
            out[o + ff + "_ln_w"] = f32(sd[p + ff + ".fn.norm.weight"])
            out[o + ff + "_ln_b"] = f32(sd[p + ff + ".fn.norm.bias"])
            out[o + ff + "_w1"] = bf16(sd[p + ff + ".fn.fn.net.0.weight"])
            out[o + ff + "_b1"] = f32(sd[p + ff + ".fn.fn.net.0.bias"])
            out[o + ff + "_w2"] = bf16(sd[p + ff + ".fn.fn.net.3.weight"])
            out[o + ff + "_b2"] = f32(sd[p + ff + ".fn.fn.net.3.bias"])
        out[o + "attn_ln_w"] = f32(sd[p + "attn.norm.weight"])
        out[o + "attn_ln_b"] = f32(sd[p + "attn.norm.bias"])
        out[o + "wqkv"] = bf16(torch.cat([sd[p + "attn.fn.to_q.weight"], sd[p + "attn.fn.to_kv.weight"]], dim=0))
        out[o + "wo"] = bf16(sd[p + "attn.fn.to_out.weight"])
        out[o + "bo"] = f32(sd[p + "attn.fn.to_out.bias"])
        out[o + "conv_ln_w"] = f32(sd[p + "conv.net.0.weight"])
        out[o + "conv_ln_b"] = f32(sd[p + "conv.net.0.bias"])
        # pointwise conv 1 feeds a GLU (value = first half of the channels, gate = second half): interleave 32 value rows
        # with their 32 gate rows so the GEMM epilogue, which owns 64 consecutive columns per thread, can gate in place
        perm = glu_interleave(sd[p + "conv.net.2.weight"].shape[0])
        out[o + "pw1_w"] = bf16(sd[p + "conv.net.2.weight"][:, :, 0][perm])
        out[o + "pw1_b"] = f32(sd[p + "conv.net.2.bias"][perm])
        # autocast runs the depthwise conv with bf16 weights
        out[o + "dw_w"] = f32(sd[p + "conv.net.4.conv.weight"][:, 0, :].to(torch.bfloat16).float())
        out[o + "dw_b"] = f32(sd[p + "conv.net.4.conv.bias"])
        out[o + "cln_w"] = f32(sd[p + "conv.net.6.weight"].reshape(-1))
        out[o + "pw2_w"] = bf16(sd[p + "conv.net.7.weight"][:, :, 0])
        out[o + "pw2_b"] = f32(sd[p + "conv.net.7.bias"])
        out[o + "post_ln_w"] = f32(sd[p + "post_norm.weight"])
        out[o + "post_ln_b"] = f32(sd[p + "post_norm.bias"])

    n_inj = len(cfg.injection_layers)
    f64 = torch.float64
    _, _, w_out, b_out, cb = quantizer_tensors(sd, max(n_inj, 1), "acoustic_model.quantizer.", device)
    proj = torch.einsum("lcd,lkd->lkc", w_out, cb)                                  # [n_inj, codes, latent]
    w_fp = sd["acoustic_feat_proj.0.weight"].to(device, f64)
    out["sem_emb"] = f32(sd["semantic_embedding.weight"])
    out["mask_token"] = f32(sd["mask_token"].reshape(-1))
    out["feat_table"] = f32(proj[0] @ w_fp.t())
    out["feat_const"] = f32(w_fp @ b_out[0] + sd["acoustic_feat_proj.0.bias"].to(device, f64))
    # the un-folded projections, for callers that hand in DAC *features* (model.acoustic_feat_proj(x), encoder.forward(injections=...))
    out["fp_w"] = bf16(sd["acoustic_feat_proj.0.weight"])
    out["fp_b"] = f32(sd["acoustic_feat_proj.0.bias"])
    out["inj_w"] = torch.stack([bf16(sd[f"encoder.project_injection.{k}.0.weight"]) for k in range(n_inj)])
    out["inj_b"] = torch.stack([f32(sd[f"encoder.project_injection.{k}.0.bias"]) for k in range(n_inj)])
    out["fp_ln_w"] = f32(sd["acoustic_feat_proj.1.weight"])
    out["fp_ln_b"] = f32(sd["acoustic_feat_proj.1.bias"])
    codes = cb.shape[1]
    inj_table = torch.zeros(4, 4, codes, d, device=device, dtype=torch.float32)
    inj_const = torch.zeros(4, d, device=device, dtype=torch.float32)
    inj_ln_w = torch.ones(4, d, device=device, dtype=torch.float32)
    inj_ln_b = torch.zeros(4, d, device=device, dtype=torch.float32)
    for k in range(n_inj):
        w_pi = sd[f"encoder.project_injection.{k}.0.weight"].to(device, f64)
        for i in range(k + 1):
            inj_table[k, i] = (proj[i] @ w_pi.t()).float()
        inj_const[k] = (w_pi @ b_out[: k + 1].sum(0) + sd[f"encoder.project_injection.{k}.0.bias"].to(device, f64)).float()
        inj_ln_w[k] = f32(sd[f"encoder.project_injection.{k}.1.weight"])
        inj_ln_b[k] = f32(sd[f"encoder.project_injection.{k}.1.bias"])
    out["inj_table"], out["inj_const"], out["inj_ln_w"], out["inj_ln_b"] = inj_table, inj_const, inj_ln_w, inj_ln_b
    out["tl_ln_w"] = f32(sd["encoder.to_logits.0.weight"])
    out["tl_ln_b"] = f32(sd["encoder.to_logits.0.bias"])
    w_head = sd["encoder.to_logits.1.weight"]                                        # [q, d, l] (EinMix 'q d l')
    out["head_w"] = bf16(w_head.permute(0, 2, 1).reshape(-1, d))                     # [q*l, d]: K-major B operand
    out["head_b"] = f32(sd["encoder.to_logits.1.bias"].reshape(-1))
    out["fine_w"] = bf16(sd["encoder.fine_head.0.weight"])
    out["fine_b"] = f32(sd["encoder.fine_head.0.bias"])
    out["rope_cos"], out["rope_sin"] = rope_tables(max_positions, d // cfg.heads, device)
    return out


def tf32_round(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest tf32 value (10 explicit mantissa bits, ties away from zero, like cvt.rna.tf32.f32), kept as fp32."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_rvq_weights(sd: dict, n_codebooks: int, prefix: str, device) -> dict:
    """Tables of the fused RVQ search (csrc/rvq_tc.cuh): stacked in_proj, normalised codebooks, G[i][j] cross tables, and the
    projected codebooks (incl. bias) for codes -> features."""
    w_in, b_in, w_out, b_out, cb = quantizer_tensors(sd, n_codebooks, prefix, device)
    L, cbd, latent = w_in.shape
    proj = torch.einsum("lcd,lkd->lkc", w_out, cb) + b_out[:, None, :]               # [L, codes, latent] incl. bias
    g = torch.einsum("idc,jkc->ijkd", w_in, proj)                                    # [L, L, codes, cb_dim]
    cb32 = cb.float()
    cbn = torch.nn.functional.normalize(cb32, dim=-1)                                # F.normalize in fp32, as the reference
    pad = 12 - L
    def padl(t):
        return torch.cat([t, torch.zeros(pad, *t.shape[1:], device=t.device, dtype=t.dtype)]) if pad else t
    g_full = torch.zeros(12, 12, cb.shape[1], cbd, device=device, dtype=torch.float32)
    g_full[:L, :L] = g.float()
    # tcgen05 path (csrc/rvq_tc.cuh): 3xTF32 operand splits. hi = tf32(v) (round to nearest, ties away), lo = tf32(v - hi).
    w_all = padl(w_in.float()).reshape(12 * cbd, latent).contiguous()               # [96, latent]: K-major B operand
    w_hi = tf32_round(w_all)
    w_lo = tf32_round(w_all - w_hi)
    cbn_p, x = padl(cbn), -0.5 * padl(cbn.pow(2).sum(-1))
    c_hi = tf32_round(cbn_p)
    c_lo = tf32_round(cbn_p - c_hi)
    x_hi = tf32_round(x)
    x_lo = tf32_round(x - x_hi)
    cb_packed = torch.zeros(12, cb.shape[1], 32, device=device, dtype=torch.float32)
    cb_packed[..., 0:8], cb_packed[..., 8:16], cb_packed[..., 16:24] = c_hi, c_hi, c_lo  # against [e_hi | e_lo | e_hi | 1 1 0..]
    cb_packed[..., 24], cb_packed[..., 25] = x_hi, x_lo
    return {
        "w_hi": w_hi.contiguous(), "w_lo": w_lo.contiguous(), "cb_packed": cb_packed.contiguous(),
        "b_in": padl(b_in.float()).reshape(-1).contiguous(),
        "cb_norm": padl(cbn).contiguous(),
        "cb_n2": padl(cbn.pow(2).sum(-1)).contiguous(),
        "g": g_full.contiguous(),
        "proj": padl(proj.float()).contiguous(),
        "n_levels": L,
    }
