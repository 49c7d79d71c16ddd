"""Synthetic weights / inputs for the oracle and the parity tests.

The deterministic generators live in edm_tts_b200/synthetic.py (bench.py needs them too, and product code must not
import oracle/); they are re-exported here. weight_norm_fold is the oracle's own restatement of the weight-norm
parametrisation and is deliberately not shared with the product code.
"""
from __future__ import annotations

import torch

from edm_tts_b200.synthetic import OracleConfig, make_inputs, make_quantizer_state_dict, make_state_dict  # noqa: F401


def weight_norm_fold(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch.nn.utils.parametrizations.weight_norm, dim=0: w = g * v / ||v|| with the norm over all dims but 0."""
    return v * (g / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1))))
