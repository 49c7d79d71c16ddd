"""GPU parity of the DAC conv decoder (edm_tts_b200/dac_decoder.py on csrc/dac_conv.cuh) against the fp32 oracle and the reference's
golden audio.

Tolerance: bf16 conv operands, fp32 accumulation and residual stream (the reference runs the decoder under bf16 autocast,
inference.py:33); against the fp32 oracle: relative L2 error of the waveform below 2e-2, max abs error below 6e-2 of the rms (floor 0.05)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(x, ref):
    x, ref = x.float().cpu(), ref.float().cpu()
    rms = ref.pow(2).mean().sqrt().item()
    rel = ((x - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item()
    mx = (x - ref).abs().max().item()
    print(f"rel L2 {rel:.2e}  max abs {mx:.2e}  rms {rms:.3f}")
    assert rel < 2e-2 and mx < 6e-2 * max(rms, 0.05)


def test_decoder_vs_reference_golden(golden_dir):
    from edm_tts_b200.dac_decoder import DACDecoder
    from edm_tts_b200.synthetic import make_decoder_state_dict

    g = torch.load(os.path.join(golden_dir, "dac_decoder_full.pt"))
    dec = DACDecoder(make_decoder_state_dict(g["input_channel"], g["channels"], (8, 5, 4, 2), g["weight_seed"]), g["input_channel"], g["channels"])
    z = torch.randn(g["B"], g["input_channel"], g["T"], generator=torch.Generator().manual_seed(g["z_seed"])) * 0.5
    audio = dec(z)
    assert tuple(audio.shape) == tuple(g["audio_shape"])
    _check(audio, g["audio"])


@pytest.mark.parametrize("B,T", [(1, 1), (3, 37), (2, 150)])
def test_decoder_vs_oracle(B, T):
    from edm_tts_b200.dac_decoder import DACDecoder
    from edm_tts_b200.synthetic import make_decoder_state_dict
    from oracle.dac_decoder import decoder_forward

    sd = make_decoder_state_dict(1024, 1536, (8, 5, 4, 2), 2)
    dec = DACDecoder(sd)
    z = torch.randn(B, 1024, T, generator=torch.Generator().manual_seed(11 + T)) * 0.5
    with torch.inference_mode():
        ref = decoder_forward(sd, z)
    audio = dec(z)
    assert audio.shape == ref.shape
    _check(audio, ref)
    assert torch.equal(audio, dec(z))
    if B > 1:      # utterances are independent; chunking the batch changes nothing
        assert torch.equal(DACDecoder(sd, max_chunk_samples=1)(z), audio)


def test_decode_from_codes_and_errors():
    from edm_tts_b200.dac import DAC
    from edm_tts_b200.synthetic import make_dac_state_dict
    from oracle import rvq as orvq  # noqa: F401
    from oracle.dac_decoder import decoder_forward

    sd = make_dac_state_dict(3)
    dac = DAC(sd)
    codes = torch.randint(0, 1024, (2, 12, 60), generator=torch.Generator().manual_seed(5))
    audio = dac.decode_from_codes(codes.cuda())
    feats = dac.codes_to_features(codes.cuda()).cpu()
    with torch.inference_mode():
        ref = decoder_forward(sd, feats, prefix="decoder.")
    _check(audio, ref)
    assert dac.decode_from_codes(codes.cuda(), length=1000).shape == (2, 1, 1000)
    with pytest.raises(ValueError):
        dac.decoder(torch.zeros(1, 512, 4))


def test_decoder_full_size_properties():
    """The S2A bench shape (utterances of 500 frames -> 160 016 samples): determinism, batch independence, bounded output and
    time-locality (the receptive field of the stack is finite: editing one latent frame changes only nearby samples)."""
    from edm_tts_b200.dac_decoder import DACDecoder
    from edm_tts_b200.synthetic import make_decoder_state_dict

    dec = DACDecoder(make_decoder_state_dict(1024, 1536, (8, 5, 4, 2), 0))
    z = torch.randn(16, 1024, 500, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6)) * 0.5
    audio = dec(z)
    assert audio.shape == (16, 1, 160016) and torch.isfinite(audio).all() and audio.abs().max() <= 1.0
    assert torch.equal(audio, dec(z))
    assert torch.equal(audio[5:7], dec(z[5:7].contiguous()))
    z2 = z.clone()
    z2[:, :, 250] = 0.0
    a2 = dec(z2)
    centre = 250 * 320
    assert not torch.equal(a2[:, :, centre - 200:centre + 200], audio[:, :, centre - 200:centre + 200])
    assert torch.equal(a2[:, :, :centre - 16000], audio[:, :, :centre - 16000]) and torch.equal(a2[:, :, centre + 16000:], audio[:, :, centre + 16000:])
    # a prefix of the latent gives the prefix of the waveform away from the cut
    ap = dec(z[:, :, :200].contiguous())
    assert torch.equal(ap[:, :, :200 * 320 - 16000], audio[:, :, :200 * 320 - 16000])


def test_pipeline_semantic_tokens_to_waveform():
    """inference.py:43-49 on the CUDA path: S2A infer_special -> DAC.decode_from_codes, against the oracle decoder fed with the same
    codes (the S2A stage itself is graded in tests/test_gpu_s2a.py)."""
    from edm_tts_b200.dac import DAC
    from edm_tts_b200.synthetic import make_dac_state_dict
    from oracle.dac_decoder import decoder_forward
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    dsd = make_dac_state_dict(0)
    # the vocoder must use the quantizer the S2A model was built with: same synthetic quantizer weights (seed 0) on both sides
    for k in list(dsd):
        if k.startswith("quantizer."):
            dsd[k] = sd["acoustic_model." + k].cpu()
    dac = DAC(dsd)
    inp = make_inputs(2, 50, 10, 4, cfg, seed=21)
    codes = model.infer_special(inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"], steps=4, seed=3)
    wav = dac.decode_from_codes(codes)
    assert wav.shape == (2, 1, 50 * 320 + 16) and torch.isfinite(wav).all()
    feats = dac.codes_to_features(codes).cpu()
    assert torch.allclose(feats, model.acoustic_model.codes_to_features(codes).cpu(), rtol=1e-5, atol=1e-5)
    with torch.inference_mode():
        ref = decoder_forward(dsd, feats, prefix="decoder.")
    _check(wav, ref)
