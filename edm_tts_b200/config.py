"""Configuration mirror of the reference's InjectionConformerConfig / DACConfig / TextToSemanticWLenConfig.

Reads the same HF-style config.json (edm_tts/models/injection_conformer/configuration.py:4-65,
edm_tts/models/dac/configuration.py:6-20) and also accepts the reference's own config objects (duck-typed).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field


@dataclass
class DACConfig:
    encoder_dim: int = 64
    encoder_rates: tuple = (2, 4, 5, 8)
    n_codebooks: int = 12
    codebook_size: int = 1024
    codebook_dim: int = 8
    sample_rate: int = 16000
    decoder_dim: int = 1536
    decoder_rates: tuple = (8, 5, 4, 2)

    @property
    def latent_dim(self) -> int:
        # Encoder doubles the channel count at every stage (dac/encoder.py:32-58)
        return self.encoder_dim * (2 ** len(self.encoder_rates))

    @classmethod
    def from_any(cls, obj) -> "DACConfig":
        if obj is None:
            return cls()
        if isinstance(obj, cls):
            return obj
        get = (lambda k, d: obj.get(k, d)) if isinstance(obj, dict) else (lambda k, d: getattr(obj, k, d))
        return cls(encoder_dim=get("encoder_dim", 64), encoder_rates=tuple(get("encoder_rates", (2, 4, 5, 8))),
                   n_codebooks=get("n_codebooks", 12), codebook_size=get("codebook_size", 1024),
                   codebook_dim=get("codebook_dim", 8), sample_rate=get("sample_rate", 16000), decoder_dim=get("decoder_dim", 1536),
                   decoder_rates=tuple(get("decoder_rates", (8, 5, 4, 2))))


@dataclass
class InjectionConformerConfig:
    hidden_size: int = 1024
    num_semantic_tokens: int = 1024
    acoustic_model_path: str = "exp/edm_tts/dac/best_model"
    encoder_config: dict = field(default_factory=lambda: dict(depth=16, heads=16, ff_mult=4, conv_kernel_size=5, dim_head=64))
    injection_layers: tuple = (4, 7, 10, 13)
    residual: bool = True
    use_injection: bool = True
    loss_all: bool = False
    dac: DACConfig = field(default_factory=DACConfig)

    @property
    def depth(self):
        return int(self.encoder_config.get("depth", 16))

    @property
    def heads(self):
        return int(self.encoder_config.get("heads", 16))

    @property
    def ff_mult(self):
        return int(self.encoder_config.get("ff_mult", 4))

    @property
    def conv_kernel_size(self):
        return int(self.encoder_config.get("conv_kernel_size", 5))

    @classmethod
    def from_any(cls, obj, dac=None) -> "InjectionConformerConfig":
        if isinstance(obj, cls):
            return obj
        get = (lambda k, d: obj.get(k, d)) if isinstance(obj, dict) else (lambda k, d: getattr(obj, k, d))
        enc = dict(get("encoder_config", {}) or {})
        return cls(hidden_size=get("hidden_size", 1024), num_semantic_tokens=get("num_semantic_tokens", 1024),
                   acoustic_model_path=get("acoustic_model_path", ""), encoder_config=enc,
                   injection_layers=tuple(get("injection_layers", (4, 7, 10, 13))), residual=bool(get("residual", True)),
                   use_injection=bool(get("use_injection", True)), loss_all=bool(get("loss_all", False)), dac=DACConfig.from_any(dac))

    @classmethod
    def from_pretrained(cls, path: str) -> "InjectionConformerConfig":
        with open(os.path.join(path, "config.json")) as f:
            raw = json.load(f)
        dac = None
        dac_path = os.path.join(raw.get("acoustic_model_path", ""), "config.json")
        if os.path.exists(dac_path):
            with open(dac_path) as f:
                dac = json.load(f)
        return cls.from_any(raw, dac)


@dataclass
class TextToSemanticWLenConfig:
    """edm_tts/models/text_to_semantic/configuration.py:4-86 (the fields the decode path reads)."""
    hidden_size: int = 512
    semantic_vocab_size: int = 1024
    text_vocab_size: int = 256
    main_encoder_args: dict = field(default_factory=lambda: dict(depth=8, heads=16, ff_mult=4, conv_kernel_size=5))
    length_predictor_args: dict = field(default_factory=lambda: dict(depth=4, heads=16, ff_mult=4, conv_kernel_size=5))
    special_tokens: dict = field(default_factory=lambda: dict(pad=0, text=1, speech=2, sep=3, mask=4))

    @classmethod
    def from_any(cls, obj) -> "TextToSemanticWLenConfig":
        if isinstance(obj, cls):
            return obj
        if obj is None:
            return cls()
        get = (lambda k, d: obj.get(k, d)) if isinstance(obj, dict) else (lambda k, d: getattr(obj, k, d))
        base = cls()
        main = dict(base.main_encoder_args, **(get("main_encoder_args", None) or {}))
        lp = dict(base.length_predictor_args, **(get("length_predictor_args", None) or {}))
        if isinstance(obj, dict):                      # constructor-style keys of configuration.py (main_encoder_num_heads, ...)
            for short, key in (("heads", "num_heads"), ("depth", "num_layers"), ("ff_mult", "ff_mult"), ("conv_kernel_size", "conv_kernel_size")):
                if f"main_encoder_{key}" in obj and "main_encoder_args" not in obj:
                    main[short] = obj[f"main_encoder_{key}"]
                if f"length_predictor_{key}" in obj and "length_predictor_args" not in obj:
                    lp[short] = obj[f"length_predictor_{key}"]
        return cls(hidden_size=get("hidden_size", 512), semantic_vocab_size=get("semantic_vocab_size", 1024), text_vocab_size=get("text_vocab_size", 256),
                   main_encoder_args=main, length_predictor_args=lp, special_tokens=dict(get("special_tokens", base.special_tokens)))

    @classmethod
    def from_pretrained(cls, path: str) -> "TextToSemanticWLenConfig":
        with open(os.path.join(path, "config.json")) as f:
            return cls.from_any(json.load(f))
