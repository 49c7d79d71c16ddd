"""Pins the oracle (oracle/) to the golden vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from oracle import rvq as orvq
from oracle import s2a as os2a
from oracle.weights import OracleConfig, make_inputs, make_quantizer_state_dict, make_state_dict

CONFIGS = {
    "small": OracleConfig(hidden=128, heads=2, depth=6, injection_layers=(1, 2, 3, 4), num_semantic=50, n_codebooks=12,
                          codebook_size=64, codebook_dim=8, latent_dim=128),
    "full": OracleConfig(),
}
_SD = {}


def state_dict(cfg_name, seed):
    key = (cfg_name, seed)
    if key not in _SD:
        _SD[key] = make_state_dict(CONFIGS[cfg_name], seed)
    return _SD[key]


S2A_CASES = ["small_s1", "small_s4", "small_s8_prompt", "full_s1", "full_s8", "full_s4_prompt"]


@pytest.mark.parametrize("name", S2A_CASES)
def test_s2a_oracle_matches_reference(name, golden_dir):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = torch.load(os.path.join(golden_dir, f"s2a_{name}.pt"))
    cfg = CONFIGS[g["cfg_name"]]
    sd = state_dict(g["cfg_name"], g["weight_seed"])
    inp = make_inputs(g["B"], g["T"], g["P"], g["steps"], cfg, seed=g["input_seed"])
    trace = {}
    with torch.inference_mode():
        codes = os2a.infer_special(sd, cfg, inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"],
                                   steps=g["steps"], temperature=g["temperature"], cat_gumbel=inp["cat_gumbel"],
                                   remask_gumbel=inp["remask_gumbel"], mode="fp32", trace=trace)
    # decisions are bit-exact: the restatement runs the same fp32 ops in the same order
    assert torch.equal(codes.to(torch.int16), g["codes"])
    for s, m in enumerate(g["step_masks"]):
        assert torch.equal(trace["step_masks"][s], m), f"mask after step {s}"
    for s, am in enumerate(g["step_argmax"]):
        assert torch.equal(trace["step_logits"][s].argmax(-1).to(torch.int16), am), f"first-level argmax step {s}"
    for s, rows in enumerate(g["step_logit_rows"]):
        torch.testing.assert_close(trace["step_logits"][s][:, g["row_idx"]], rows, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(trace["all_logits"][:, :, g["final_row_idx"]], g["final_logit_rows"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["small_s8_prompt", "full_s8"])
def test_s2a_oracle_bf16_mode_tracks_reference_autocast(name, golden_dir):
    """mode='bf16' states the rounding points the CUDA kernels use; it must stay within bf16 noise of the reference run
    under torch.autocast('cpu', bfloat16) (first-level logits of step 0, no cascade)."""
    g = torch.load(os.path.join(golden_dir, f"s2a_{name}.pt"))
    cfg = CONFIGS[g["cfg_name"]]
    sd = state_dict(g["cfg_name"], g["weight_seed"])
    inp = make_inputs(g["B"], g["T"], g["P"], g["steps"], cfg, seed=g["input_seed"])
    with torch.inference_mode():
        x, _, _, P = os2a.build_encoder_input(sd, cfg, inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"], "bf16")
        logits = os2a.forward_first_level(sd, cfg, x, P, "bf16")
    diff = (logits[:, g["row_idx"]] - g["bf16_step0_logit_rows"]).abs()
    assert diff.max().item() < 0.12 and diff.mean().item() < 0.02, (diff.max().item(), diff.mean().item())


@pytest.mark.parametrize("name", ["small", "full"])
def test_rvq_oracle_matches_reference(name, golden_dir):
    g = torch.load(os.path.join(golden_dir, f"rvq_{name}.pt"))
    cfg = CONFIGS[g["cfg_name"]]
    sd = make_quantizer_state_dict(cfg, g["weight_seed"])
    z = torch.randn(g["B"], cfg.latent_dim, g["T"], generator=torch.Generator().manual_seed(g["z_seed"]))
    with torch.inference_mode():
        out = orvq.rvq_forward(sd, cfg, z)
    assert torch.equal(out["codes"].to(torch.int16), g["codes"])
    torch.testing.assert_close(out["latents"][:, :, :8], g["latents_head"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["z"][:, :16, :8], g["zq_head"], rtol=1e-5, atol=1e-5)
    assert abs(out["z"].double().sum().item() - g["zq_sum"]) < 1e-2
    # from_codes / from_codes_unreduced restatement (used by the S2A path)
    full_sd = {"acoustic_model.quantizer." + k: v for k, v in sd.items()}
    codes = g["codes"].long()
    feats = os2a.codes_to_features(full_sd, cfg, codes)
    torch.testing.assert_close(feats[:, :16, :8], g["feats_head"], rtol=1e-5, atol=1e-5)
    unred = os2a.codes_to_features_unreduced(full_sd, cfg, codes[:, :4])
    torch.testing.assert_close(unred[:, :, :16, :8], g["unred_head"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["small", "full", "small_loss_all"])
def test_training_forward_oracle_matches_reference(name, golden_dir):
    """InjectionConformerModel.forward in eval mode (loss + arg-max codes) with the mask draw injected; the loss_all case pins
    the un-flattened [b, q, t] shape of the returned codes."""
    g = torch.load(os.path.join(golden_dir, f"train_fwd_{name}.pt"))
    cfg = CONFIGS[g["cfg_name"]]
    sd = state_dict(g["cfg_name"], g["weight_seed"])
    with torch.inference_mode():
        out = os2a.training_forward(sd, cfg, g["acoustic_tokens"].long(), g["semantic_tokens"].long(), g["mask"], loss_all=g.get("loss_all", False))
    assert abs(out["loss"].item() - g["loss"]) < 1e-4, (out["loss"].item(), g["loss"])
    assert out["output_acoustic_codes"].shape == g["output_codes"].shape
    assert torch.equal(out["output_acoustic_codes"].to(torch.int16), g["output_codes"])
    torch.testing.assert_close(out["logits"][:, :, g["row_idx"]], g["logit_rows"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["small", "full"])
def test_dac_encoder_oracle_matches_reference(golden_dir, name):
    """oracle/dac_encoder.py against z of the unmodified reference Encoder (fp32 CPU): same conv algorithm -> 1e-5."""
    from edm_tts_b200.synthetic import make_encoder_state_dict
    from oracle.dac_encoder import encoder_forward

    g = torch.load(os.path.join(golden_dir, f"dac_encoder_{name}.pt"))
    sd = make_encoder_state_dict(g["encoder_dim"], (2, 4, 5, 8), g["weight_seed"])
    audio = (torch.randn(g["B"], 1, g["L"], generator=torch.Generator().manual_seed(g["audio_seed"])) * 0.3).clamp(-1, 1)
    with torch.inference_mode():
        z, stages = encoder_forward(sd, audio, return_stages=True)
    assert tuple(z.shape) == tuple(g["z_shape"])
    print("stage rms:", [round(s.pow(2).mean().sqrt().item(), 3) for s in stages])
    if g["z"] is not None:
        torch.testing.assert_close(z, g["z"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(z[:, :32, :16], g["z_head"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(z[:, -32:, -16:], g["z_tail"], rtol=1e-5, atol=1e-5)
    assert abs(z.double().sum().item() - g["z_sum"]) < 1e-3 * max(1.0, g["z_abs_sum"] * 1e-3)


@pytest.mark.parametrize("name", ["small", "full"])
def test_dac_decoder_oracle_matches_reference(golden_dir, name):
    """oracle/dac_decoder.py against the audio of the unmodified reference Decoder (fp32 CPU)."""
    from edm_tts_b200.synthetic import make_decoder_state_dict
    from oracle.dac_decoder import decoder_forward

    g = torch.load(os.path.join(golden_dir, f"dac_decoder_{name}.pt"))
    sd = make_decoder_state_dict(g["input_channel"], g["channels"], (8, 5, 4, 2), g["weight_seed"])
    z = torch.randn(g["B"], g["input_channel"], g["T"], generator=torch.Generator().manual_seed(g["z_seed"])) * 0.5
    with torch.inference_mode():
        audio, stages = decoder_forward(sd, z, return_stages=True)
    assert tuple(audio.shape) == tuple(g["audio_shape"])
    print("stage rms:", [round(s.pow(2).mean().sqrt().item(), 3) for s in stages])
    if g["audio"] is not None:
        torch.testing.assert_close(audio, g["audio"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(audio[:, :, :256], g["audio_head"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(audio[:, :, -256:], g["audio_tail"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(audio[:, :, ::37], g["audio_strided"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["small_s4", "small_s1", "base_s3", "train_s4"])
def test_t2s_oracle_matches_reference(golden_dir, name):
    """oracle/t2s.py against TextToSemanticWLen.infer of the unmodified reference (fp32 CPU, injected noise): predicted length,
    per-iteration masks and final tokens bit-exact, logit rows to 1e-4. Covers hidden 128 / 512 (dh 32) / 384 (dh 48)."""
    from edm_tts_b200.synthetic import T2SConfig, make_t2s_noise, make_t2s_state_dict
    from oracle import t2s as ot2s
    from tests.golden.make_golden_cfg import T2S_CONFIGS

    g = torch.load(os.path.join(golden_dir, f"t2s_{name}.pt"))
    cfg = T2SConfig(**T2S_CONFIGS[g["cfg_name"]])
    sd = make_t2s_state_dict(cfg, g["weight_seed"])
    with torch.inference_mode():
        if g["gt_length"] is None:
            length, raw = ot2s.predict_length(sd, cfg, ot2s.text_tokens_of(g["text"], cfg), return_raw=True)
            assert int(length) == g["length"]
            assert abs(raw.item() - g["raw_log_length"]) < 1e-4
        L = len(g["text"].encode("utf-8")) + g["length"] + 4
        noise = make_t2s_noise(L, g["pred_iters"], cfg, seed=g["noise_seed"])
        tr = {}
        tokens = ot2s.infer(sd, cfg, g["text"], pred_iters=g["pred_iters"], gt_length=g["gt_length"], cat_gumbel=noise["cat_gumbel"],
                            remask_gumbel=noise["remask_gumbel"], trace=tr)
    assert torch.equal(tokens.to(torch.int16), g["tokens"])
    if g["masks"] is not None:
        assert torch.equal(torch.stack(tr["step_masks"]), g["masks"])
    torch.testing.assert_close(tr["step_logits"][0][0][g["row_idx"]], g["first_rows"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(tr["step_logits"][-1][0][g["row_idx"]], g["last_rows"], rtol=1e-4, atol=1e-4)
