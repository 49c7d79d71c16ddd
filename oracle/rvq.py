"""Restatement of DAC ResidualVectorQuantize.forward in eval mode (test infrastructure; see oracle/__init__.py).

Reference: edm_tts/models/dac/vector_quantizer.py:146-210 (residual loop), :33-67 (VectorQuantize.forward),
:75-91 (decode_latents: L2-normalised nearest code, first arg-max wins).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .weights import OracleConfig, weight_norm_fold


def quantizer_weights(sd, cfg: OracleConfig, prefix=""):
    out = []
    for i in range(cfg.n_codebooks):
        q = f"{prefix}quantizers.{i}."
        w_in = weight_norm_fold(sd[q + "in_proj.parametrizations.weight.original0"], sd[q + "in_proj.parametrizations.weight.original1"])[:, :, 0]
        w_out = weight_norm_fold(sd[q + "out_proj.parametrizations.weight.original0"], sd[q + "out_proj.parametrizations.weight.original1"])[:, :, 0]
        out.append(dict(w_in=w_in, b_in=sd[q + "in_proj.bias"], w_out=w_out, b_out=sd[q + "out_proj.bias"], codebook=sd[q + "codebook.weight"]))
    return out


def rvq_forward(sd, cfg: OracleConfig, z: torch.Tensor, n_quantizers=None, prefix="", forced_codes=None, return_margins=False):
    """z [B, latent, T] -> dict(codes [B, L, T], z [B, latent, T], latents [B, L*cb_dim, T]).
    forced_codes teacher-forces the residual update (parity protocol). margins = top1 - top2 of (-dist) per decision."""
    qs = quantizer_weights(sd, cfg, prefix)
    dev = z.device
    n_q = n_quantizers or cfg.n_codebooks
    residual, z_q = z, 0
    codes, latents, margins = [], [], []
    for i, q in enumerate(qs):
        w_in, b_in, w_out, b_out, cb = (q[k].to(dev) for k in ("w_in", "b_in", "w_out", "b_out", "codebook"))
        z_e = F.conv1d(residual, w_in[:, :, None], b_in)                               # in_proj :57
        enc = z_e.permute(0, 2, 1).reshape(-1, z_e.shape[1])
        enc_n, cb_n = F.normalize(enc), F.normalize(cb)
        dist = enc_n.pow(2).sum(1, keepdim=True) - 2 * enc_n @ cb_n.t() + cb_n.pow(2).sum(1, keepdim=True).t()
        neg = -dist
        idx = neg.max(1)[1].view(z.shape[0], -1)
        if return_margins:
            top2 = neg.topk(2, dim=1)[0]
            margins.append((top2[:, 0] - top2[:, 1]).view(z.shape[0], -1))
        codes.append(idx)
        latents.append(z_e)
        use = forced_codes[:, i] if forced_codes is not None else idx
        z_q_i = F.conv1d(F.embedding(use, cb).transpose(1, 2), w_out[:, :, None], b_out)  # decode_code + out_proj :65
        if i < n_q + 1:  # quantizer-dropout mask in eval mode: i < n_quantizers + 1 (:183,193)
            z_q = z_q + z_q_i
        residual = residual - z_q_i
    out = {"codes": torch.stack(codes, dim=1), "z": z_q, "latents": torch.cat(latents, dim=1)}
    if return_margins:
        out["margins"] = torch.stack(margins, dim=1)
    return out
