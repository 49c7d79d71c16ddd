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
