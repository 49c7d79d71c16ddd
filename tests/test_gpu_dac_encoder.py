"""GPU parity of the DAC conv encoder (csrc/dac_conv.cuh) against the fp32 oracle and the reference's golden z.

Tolerance: the kernels use bf16 conv operands with fp32 accumulation and an fp32 residual stream (the reference itself runs this
stack under bf16 autocast, dump_tokens.py:213); against the fp32 oracle that is ~2^-9 relative per operand, averaged over K >= 448
products and ~27 convs -> relative L2 error of z below 1e-2, max abs error below 4e-2 of the rms."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _audio(B, L):
    return (torch.randn(B, 1, L, generator=torch.Generator().manual_seed(7 + L)) * 0.3).clamp(-1, 1)


def _check(z, ref):
    z, ref = z.float().cpu(), ref.float().cpu()
    rms = ref.pow(2).mean().sqrt().item()
    rel = ((z - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item()
    mx = (z - ref).abs().max().item()
    print(f"rel L2 {rel:.2e}  max abs {mx:.2e}  rms {rms:.3f}")
    assert rel < 1e-2 and mx < 4e-2 * max(rms, 0.1)


def test_encoder_vs_reference_golden(golden_dir):
    from edm_tts_b200.dac_encoder import DACEncoder
    from edm_tts_b200.synthetic import make_encoder_state_dict

    g = torch.load(os.path.join(golden_dir, "dac_encoder_full.pt"))
    enc = DACEncoder(make_encoder_state_dict(g["encoder_dim"], (2, 4, 5, 8), g["weight_seed"]), g["encoder_dim"])
    z = enc(_audio(g["B"], g["L"]), out_dtype=torch.float32)
    assert tuple(z.shape) == tuple(g["z_shape"])
    _check(z, g["z"])


@pytest.mark.parametrize("B,L", [(1, 320), (3, 16000 + 77), (2, 48000)])
def test_encoder_vs_oracle(B, L):
    from edm_tts_b200.dac_encoder import DACEncoder
    from edm_tts_b200.synthetic import make_encoder_state_dict
    from oracle.dac_encoder import encoder_forward

    sd = make_encoder_state_dict(64, (2, 4, 5, 8), 1)
    enc = DACEncoder(sd, 64)
    audio = _audio(B, L)
    with torch.inference_mode():
        ref = encoder_forward(sd, audio)
    z32 = enc(audio, out_dtype=torch.float32)
    assert z32.shape == ref.shape
    _check(z32, ref)
    zb = enc(audio)                                   # bf16 out = the same values rounded once
    assert zb.dtype == torch.bfloat16 and torch.equal(zb, z32.to(torch.bfloat16))
    # the fused ResidualUnit kernel performs the same arithmetic in the same order as the two-launch form
    assert torch.equal(DACEncoder(sd, 64, fused=False)(audio, out_dtype=torch.float32), z32)
    # sequences are independent and chunking the batch changes nothing
    if B > 1:
        enc2 = DACEncoder(sd, 64, max_chunk_samples=L)
        assert torch.equal(enc2(audio, out_dtype=torch.float32), z32)


def test_encoder_errors():
    from edm_tts_b200.dac_encoder import DACEncoder
    from edm_tts_b200.synthetic import make_encoder_state_dict

    enc = DACEncoder(make_encoder_state_dict(64, (2, 4, 5, 8), 0), 64)
    with pytest.raises(ValueError):
        enc(torch.zeros(2, 2, 1000))
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 1, 100))
    with pytest.raises(ValueError):
        DACEncoder(make_encoder_state_dict(8, (2, 4, 5, 8), 0), 8)


def test_dac_encode_to_codes_vs_oracle():
    """audio -> codes through DAC.encode_to_codes. The encoder's bf16-operand error (rel 6e-3 of z) moves nearest-code decisions
    that sit close to a boundary, so agreement with the fp32 oracle is statistical at the first level (the reference's own bf16
    autocast path has the same property); given the SAME z the RVQ stage itself is bit-exact (tests/test_gpu_rvq.py), which is
    re-checked here by feeding the kernel's own z to the oracle quantizer."""
    from edm_tts_b200.dac import DAC
    from edm_tts_b200.synthetic import make_dac_state_dict
    from oracle import rvq as orvq
    from oracle.dac_encoder import encoder_forward
    from oracle.weights import OracleConfig

    sd = make_dac_state_dict(3)
    dac = DAC(sd)
    audio = _audio(2, 32000)
    codes = dac.encode_to_codes(audio).cpu()
    assert codes.shape == (2, 12, 100) and codes.dtype == torch.int64
    with torch.inference_mode():
        z_ref = encoder_forward(sd, audio, prefix="encoder.")
        ref = orvq.rvq_forward(sd, OracleConfig(), z_ref, prefix="quantizer.")["codes"]
    agree0 = (codes[:, 0] == ref[:, 0]).float().mean().item()
    print("first-level agreement with the fp32 oracle:", agree0)
    assert agree0 > 0.9
    # same z -> same codes (bit-exact except near-ties)
    z = dac.encoder(audio, out_dtype=torch.bfloat16)
    with torch.inference_mode():
        own = orvq.rvq_forward(sd, OracleConfig(), z.float().cpu(), prefix="quantizer.", return_margins=True)
    mism = codes != own["codes"]
    clean = ~(mism.cumsum(1) > 0)
    assert (own["margins"][mism & (mism.cumsum(1) == 1)] < 2e-5).all()
    assert torch.equal(codes[clean], own["codes"][clean]) and mism[:, 0].float().mean().item() < 1e-2
    out = dac.encode(audio)
    assert set(out) >= {"z", "codes", "latents", "length"} and out["z"].shape == (2, 1024, 100)
    audio_hat = dac.decode_from_codes(codes, length=32000)
    assert audio_hat.shape == (2, 1, 32000) and torch.isfinite(audio_hat).all()


def test_encoder_full_size_config4_properties():
    """dump_tokens shape (60 s segments, L = 960160 -> T = 3000): determinism, batch independence, and time-locality: frames far
    from an edit of the audio are unchanged (the receptive field of the stack is finite)."""
    from edm_tts_b200.dac_encoder import DACEncoder
    from edm_tts_b200.synthetic import make_encoder_state_dict

    enc = DACEncoder(make_encoder_state_dict(64, (2, 4, 5, 8), 0), 64)
    L_in = 960160
    audio = (torch.randn(4, 1, L_in, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.3).clamp(-1, 1)
    z = enc(audio)
    assert z.shape == (4, 1024, 3000) and torch.isfinite(z.float()).all()
    assert torch.equal(z, enc(audio))
    assert torch.equal(z[2:3], enc(audio[2:3].contiguous()))
    a2 = audio.clone()
    a2[:, :, 480000:480320] = 0.0
    z2 = enc(a2)
    assert not torch.equal(z2[:, :, 1495:1505], z[:, :, 1495:1505])
    assert torch.equal(z2[:, :, :1400], z[:, :, :1400]) and torch.equal(z2[:, :, 1600:], z[:, :, 1600:])
    # a prefix of the audio gives the prefix of z away from the cut
    zp = enc(audio[:, :, :320000].contiguous())
    assert torch.equal(zp[:, :, :900], z[:, :, :900])
