"""Restatement of the DAC convolutional encoder (test infrastructure; see oracle/__init__.py).

Reference: edm_tts/models/dac/encoder.py:11-58 (Encoder / EncoderBlock), edm_tts/models/dac/nn_layers.py:8-47
(WNConv1d = weight_norm(Conv1d), Snake1d: x + 1/(alpha + 1e-9) * sin(alpha x)^2, ResidualUnit: x + conv1(snake(conv7_dilated(snake(x))))),
caller edm_tts/models/dac/modeling_dac.py:163-167 (encode_to_codes: z = encoder(audio); codes = quantizer(z)["codes"]).
Pinned to the unmodified reference by tests/golden/dac_encoder_*.pt (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .weights import weight_norm_fold


def snake(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)     # nn_layers.py:24-29


def _conv(sd, key, x, **kw):
    w = weight_norm_fold(sd[key + ".parametrizations.weight.original0"], sd[key + ".parametrizations.weight.original1"]).to(x.device)
    return F.conv1d(x, w, sd[key + ".bias"].to(x.device), **kw)


def encoder_forward(sd, audio: torch.Tensor, rates=(2, 4, 5, 8), prefix: str = "", return_stages: bool = False):
    """audio [B, 1, L] -> z [B, encoder_dim * 2^len(rates), T]. return_stages also gives the output of every top-level block."""
    dev = audio.device
    stages = []
    x = _conv(sd, f"{prefix}block.0", audio, padding=3)                       # encoder.py:38
    stages.append(x)
    n = 1
    for stride in rates:
        for u, dilation in enumerate((1, 3, 9)):                              # encoder.py:15-17
            ru = f"{prefix}block.{n}.block.{u}.block."
            h = snake(x, sd[ru + "0.alpha"].to(dev))
            h = _conv(sd, ru + "1", h, dilation=dilation, padding=3 * dilation)   # nn_layers.py:36-40: pad = (7 - 1) * dilation // 2
            h = snake(h, sd[ru + "2.alpha"].to(dev))
            h = _conv(sd, ru + "3", h)
            x = x + h                                                          # nn_layers.py:46
        x = snake(x, sd[f"{prefix}block.{n}.block.3.alpha"].to(dev))
        x = _conv(sd, f"{prefix}block.{n}.block.4", x, stride=stride, padding=math.ceil(stride / 2))   # encoder.py:19-25
        stages.append(x)
        n += 1
    x = snake(x, sd[f"{prefix}block.{n}.alpha"].to(dev))
    x = _conv(sd, f"{prefix}block.{n + 1}", x, padding=1)                      # encoder.py:47-50
    return (x, stages) if return_stages else x
