"""Restatement of the DAC convolutional decoder (test infrastructure; see oracle/__init__.py).

Reference: edm_tts/models/dac/decoder.py:11-62 (DecoderBlock: Snake, weight-normed ConvTranspose1d(kernel 2s, stride s,
padding floor(s/2), output_padding s % 2), three dilated ResidualUnits; Decoder: first conv k=7, the blocks, Snake, conv k=7 to one
channel, tanh), edm_tts/models/dac/nn_layers.py:8-47, caller DAC.decode / decode_from_codes (modeling_dac.py:141-171).
Pinned to the unmodified reference by tests/golden/dac_decoder_*.pt (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .dac_encoder import snake
from .weights import weight_norm_fold


def _w(sd, key, dev):
    return weight_norm_fold(sd[key + ".parametrizations.weight.original0"], sd[key + ".parametrizations.weight.original1"]).to(dev), sd[key + ".bias"].to(dev)


def decoder_forward(sd, z: torch.Tensor, rates=(8, 5, 4, 2), prefix: str = "", return_stages: bool = False):
    """z [B, D, T] -> audio [B, 1, L]. return_stages also gives the activation after the first conv and after every block."""
    dev = z.device
    stages = []
    x = F.conv1d(z, *_w(sd, f"{prefix}model.0", dev), padding=3)                                   # decoder.py:45
    stages.append(x)
    n = 1
    for stride in rates:
        blk = f"{prefix}model.{n}.block."
        x = snake(x, sd[blk + "0.alpha"].to(dev))
        w, b = _w(sd, blk + "1", dev)          # weight_norm (dim 0) of ConvTranspose1d normalises per INPUT channel
        x = F.conv_transpose1d(x, w, b, stride=stride, padding=stride // 2, output_padding=stride % 2)   # decoder.py:15-23
        for u, dilation in enumerate((1, 3, 9)):                                                   # decoder.py:24-26
            ru = f"{blk}{2 + u}.block."
            h = snake(x, sd[ru + "0.alpha"].to(dev))
            h = F.conv1d(h, *_w(sd, ru + "1", dev), dilation=dilation, padding=3 * dilation)
            h = snake(h, sd[ru + "2.alpha"].to(dev))
            h = F.conv1d(h, *_w(sd, ru + "3", dev))
            x = x + h
        stages.append(x)
        n += 1
    x = snake(x, sd[f"{prefix}model.{n}.alpha"].to(dev))
    x = torch.tanh(F.conv1d(x, *_w(sd, f"{prefix}model.{n + 1}", dev), padding=3))                 # decoder.py:55-59
    return (x, stages) if return_stages else x
