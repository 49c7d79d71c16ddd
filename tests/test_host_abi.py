"""CPU checks of the host side: the C-ABI library loads and exports every symbol include/edm_s2a.h declares, the config
mirror reads the reference's config.json format, and the product path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "edm_s2a.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(edm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "edm_tts_b200", "libedm_s2a.so"))
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/edm_s2a.h but not exported"
    from edm_tts_b200 import _lib

    assert set(_lib.EXPORTED_SYMBOLS) == set(syms), set(_lib.EXPORTED_SYMBOLS) ^ set(syms)
    assert _lib.lib().edm_abi_version() == 1


def test_weight_name_table_matches_packer():
    """Every weight the library asks for is produced by the packer's naming scheme (no compute, no GPU)."""
    from edm_tts_b200 import _lib

    c = _lib.S2AConfig()
    c.hidden, c.heads, c.depth, c.ff_mult, c.conv_kernel = 1024, 16, 16, 4, 5
    c.num_quantizers, c.num_codes, c.num_semantic, c.n_injection = 12, 1024, 1024, 4
    for i, l in enumerate((4, 7, 10, 13)):
        c.injection_layers[i] = l
    c.residual, c.max_positions = 1, 4096
    lib = _lib.lib()
    n = lib.edm_s2a_num_weights(ctypes.byref(c))
    assert n == 16 * 28 + 18
    names = [lib.edm_s2a_weight_name(ctypes.byref(c), i).decode() for i in range(n)]
    assert names[0] == "blocks.0.ff1_ln_w" and names[-1] == "rope_sin" and len(set(names)) == n
    c.hidden = 512
    assert lib.edm_s2a_num_weights(ctypes.byref(c)) < 0  # unsupported dims are refused, not emulated


def test_config_mirror_reads_reference_json(tmp_path):
    from edm_tts_b200.config import InjectionConformerConfig

    ref_cfg = {"acoustic_model_path": str(tmp_path / "dac"), "encoder_config": {"conv_kernel_size": 5, "depth": 16, "dim_head": 64, "ff_mult": 4, "heads": 16},
               "hidden_size": 1024, "injection_layers": [4, 7, 10, 13], "loss_all": False, "num_semantic_tokens": 1024, "residual": True, "use_injection": True}
    (tmp_path / "dac").mkdir()
    json.dump(ref_cfg, open(tmp_path / "config.json", "w"))
    json.dump({"codebook_dim": 8, "codebook_size": 1024, "encoder_dim": 64, "encoder_rates": [2, 4, 5, 8], "n_codebooks": 12}, open(tmp_path / "dac" / "config.json", "w"))
    cfg = InjectionConformerConfig.from_pretrained(str(tmp_path))
    assert (cfg.hidden_size, cfg.depth, cfg.heads, cfg.ff_mult, cfg.conv_kernel_size) == (1024, 16, 16, 4, 5)
    assert cfg.injection_layers == (4, 7, 10, 13) and cfg.dac.latent_dim == 1024 and cfg.dac.n_codebooks == 12


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from edm_tts_b200 import InjectionConformerModel, ResidualVectorQuantize, _lib
    from edm_tts_b200.config import InjectionConformerConfig

    with pytest.raises(_lib.EdmError):
        InjectionConformerModel(InjectionConformerConfig(), {})
    with pytest.raises(_lib.EdmError):
        ResidualVectorQuantize({})


def test_weight_packing_algebra_cpu():
    """The folded lookup tables equal the reference's codes -> feature -> Linear chain (checked against the oracle, fp32)."""
    from edm_tts_b200.config import DACConfig, InjectionConformerConfig
    from edm_tts_b200.weights import pack_rvq_weights, pack_s2a_weights
    from oracle import s2a as os2a
    from oracle.weights import OracleConfig, make_state_dict

    ocfg = OracleConfig(depth=1, injection_layers=(0,), hidden=1024)
    sd = make_state_dict(ocfg, 3)
    # packer expects 4 injection tables; give it a 1-layer / 1-injection config
    hcfg = InjectionConformerConfig(encoder_config=dict(depth=1, heads=16, ff_mult=4, conv_kernel_size=5), injection_layers=(0,))
    w = pack_s2a_weights(sd, hcfg, "cpu", max_positions=64)
    codes = torch.randint(0, 1024, (2, 1, 9))
    feat = os2a.codes_to_features(sd, ocfg, codes).transpose(1, 2)
    want = torch.nn.functional.linear(feat, sd["acoustic_feat_proj.0.weight"], sd["acoustic_feat_proj.0.bias"])
    got = w["feat_table"][codes[:, 0]] + w["feat_const"]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    want = torch.nn.functional.linear(feat, sd["encoder.project_injection.0.0.weight"], sd["encoder.project_injection.0.0.bias"])
    got = w["inj_table"][0, 0][codes[:, 0]] + w["inj_const"][0]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    assert w["head_w"].shape == (12 * 1024, 1024) and w["head_w"].dtype == torch.bfloat16
    torch.testing.assert_close(w["head_w"][1024:2048].float(), sd["encoder.to_logits.1.weight"][1].t().to(torch.bfloat16).float())
    r = pack_rvq_weights({k[len("acoustic_model.quantizer."):]: v for k, v in sd.items() if k.startswith("acoustic_model.quantizer.")}, 12, "", "cpu")
    assert r["g"].shape == (12, 12, 1024, 8) and r["proj"].shape == (12, 1024, 1024)
