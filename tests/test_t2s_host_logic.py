"""CPU tests of the text-to-semantic host logic: the padded-head packing of the attention weights (edm_tts_b200/t2s.py) is an exact
re-arrangement of the reference attention, the config mirror reads the reference's config.json layouts, and the weight packer covers
every key of the reference state dict."""
import json

import pytest
import torch


@pytest.mark.parametrize("heads,dh", [(4, 32), (8, 48), (2, 64), (4, 8)])
def test_padded_heads_reproduce_rotary_attention(heads, dh):
    from edm_tts_b200.t2s import _pad_heads, _rope_tables
    from oracle import s2a as os2a

    torch.manual_seed(heads * 100 + dh)
    d, n = heads * dh, 11
    x = torch.randn(1, n, d)
    wq, wkv, wo = torch.randn(d, d) / d ** 0.5, torch.randn(2 * d, d) / d ** 0.5, torch.randn(d, d) / d ** 0.5
    # reference attention (conformer.py:128-146): split heads, rotate-half RoPE on q and k, softmax(q k^T dh^-1/2) v, merge, out-proj
    q, k, v = (t.view(1, n, heads, dh).transpose(1, 2) for t in (x @ wq.t(), x @ wkv[:d].t(), x @ wkv[d:].t()))
    freqs = os2a.rotary_freqs(n, dh, "cpu")
    ref = torch.nn.functional.scaled_dot_product_attention(os2a._rope(freqs, q), os2a._rope(freqs, k), v).transpose(1, 2).reshape(1, n, d) @ wo.t()
    # the packed form the kernels see: 64-column heads, rotary pairs (j, j + 32), tables [pos, 32]
    wqkv = torch.cat([_pad_heads(wq, heads, 0), _pad_heads(wkv[:d], heads, 0), _pad_heads(wkv[d:], heads, 0)])
    wo_p = _pad_heads(wo, heads, 1)
    assert wqkv.shape == (3 * heads * 64, d) and wo_p.shape == (d, heads * 64)
    cos, sin = _rope_tables(n, dh, "cpu")
    qkv = (x[0] @ wqkv.t()).view(n, 3, heads, 64)
    def rope(t):                                        # gemm_epilogue_rope64: lo = columns [0, 32), hi = [32, 64)
        lo, hi = t[..., :32], t[..., 32:]
        c, s = cos[:, None, :], sin[:, None, :]
        return torch.cat([lo * c - hi * s, hi * c + lo * s], dim=-1)
    qp, kp, vp = rope(qkv[:, 0]), rope(qkv[:, 1]), qkv[:, 2]
    att = torch.softmax(torch.einsum("qhd,khd->hqk", qp, kp) * dh ** -0.5, dim=-1)
    out = torch.einsum("hqk,khd->qhd", att, vp).reshape(n, heads * 64) @ wo_p.t()
    torch.testing.assert_close(out, ref[0], rtol=1e-4, atol=1e-4)


def test_t2s_config_mirror(tmp_path):
    from edm_tts_b200.config import TextToSemanticWLenConfig

    base = TextToSemanticWLenConfig.from_any(None)
    assert (base.hidden_size, base.main_encoder_args["heads"], base.main_encoder_args["depth"], base.length_predictor_args["depth"]) == (512, 16, 8, 4)
    # config.json as PretrainedConfig.save_pretrained writes it (configuration.py:53-83 stores the *_args dicts)
    raw = dict(hidden_size=384, semantic_vocab_size=1024, text_vocab_size=256,
               main_encoder_args=dict(depth=12, heads=8, ff_mult=4, conv_kernel_size=5, dim_head=48, attn_flash=True),
               length_predictor_args=dict(depth=4, heads=8, ff_mult=4, conv_kernel_size=5, dim_head=48),
               special_tokens=dict(pad=0, text=1, speech=2, sep=3, mask=4))
    (tmp_path / "config.json").write_text(json.dumps(raw))
    cfg = TextToSemanticWLenConfig.from_pretrained(str(tmp_path))
    assert (cfg.hidden_size, cfg.main_encoder_args["heads"], cfg.main_encoder_args["depth"]) == (384, 8, 12)
    # constructor-style keys (train_config.yaml's extra_model_params path)
    cfg2 = TextToSemanticWLenConfig.from_any(dict(hidden_size=384, main_encoder_num_heads=8, main_encoder_num_layers=12, length_predictor_num_heads=8))
    assert (cfg2.main_encoder_args["heads"], cfg2.main_encoder_args["depth"], cfg2.length_predictor_args["heads"]) == (8, 12, 8)


def test_pack_t2s_weights_covers_the_state_dict():
    from edm_tts_b200.config import TextToSemanticWLenConfig
    from edm_tts_b200.synthetic import T2SConfig, make_t2s_state_dict
    from edm_tts_b200.t2s import pack_t2s_weights

    cfg = T2SConfig(hidden=128, heads=4, depth=2, lp_heads=2, lp_depth=1)
    sd = make_t2s_state_dict(cfg, 0)
    hcfg = TextToSemanticWLenConfig(hidden_size=128, main_encoder_args=dict(depth=2, heads=4), length_predictor_args=dict(depth=1, heads=2))
    w = pack_t2s_weights(sd, hcfg, "cpu", 64)
    assert w["blocks.0.wqkv"].shape == (3 * 4 * 64, 128) and w["lp_blocks.0.wqkv"].shape == (3 * 2 * 64, 128)
    assert w["blocks.1.wo"].shape == (128, 256) and w["rope_cos"].shape == (64, 32) and w["emb"].shape == (cfg.total_tokens, 128)
    assert w["blocks.0.pw1_w"].shape == (512, 128) and w["blocks.0.dw_w"].shape == (256, 5)
    with pytest.raises(KeyError):
        pack_t2s_weights({k: v for k, v in sd.items() if k != "pred_head.weight"}, hcfg, "cpu", 64)
