"""Host-side algebra of the DAC conv paths, checked on CPU against torch's own convs (no kernels involved):
the transposed conv as a 2-tap conv into a shifted output view (edm_tts_b200/dac_decoder.py), and the strided conv as a 2-tap conv
over the padded operand viewed as [time / s][s * C] (edm_tts_b200/dac_encoder.py)."""
import math

import pytest
import torch
import torch.nn.functional as F


@pytest.mark.parametrize("stride", [8, 5, 4, 2])
def test_conv_transpose_as_two_tap_conv(stride):
    from edm_tts_b200.dac_decoder import conv_transpose_length, pack_conv_transpose

    g = torch.Generator().manual_seed(stride)
    c_in, c_out, cp, L_in, B = 12, 5, 8, 9, 2
    wt = torch.randn(c_in, c_out, 2 * stride, generator=g)
    bias = torch.randn(c_out, generator=g)
    x = torch.randn(B, c_in, L_in, generator=g)
    pad = stride // 2
    ref = F.conv_transpose1d(x, wt, bias, stride=stride, padding=pad, output_padding=stride % 2)
    L_out = conv_transpose_length(L_in, stride)
    assert ref.shape[-1] == L_out
    wp = pack_conv_transpose(wt, stride, cp)                                    # [stride * cp, 2 * c_in]
    xl = x.transpose(1, 2)                                                      # channel-last [B, L_in, c_in]
    xz = F.pad(xl, (0, 0, 1, 1))                                                # rows -1 and L_in read as zero
    a = torch.cat([xz[:, :-1], xz[:, 1:]], dim=-1)                              # view row q: [x[q - 1] | x[q]], q = 0 .. L_in
    view = a @ wp.t()                                                           # [B, L_in + 1, stride * cp]
    flat = view.reshape(B, (L_in + 1) * stride, cp)                             # row q * stride + r  <->  output time q * stride + r - pad
    out = flat[:, pad:pad + L_out, :c_out] + bias
    torch.testing.assert_close(out.transpose(1, 2), ref, rtol=1e-5, atol=1e-5)
    assert flat[:, :, c_out:].abs().max() == 0                                  # padded channels stay zero


@pytest.mark.parametrize("stride", [2, 4, 5, 8])
def test_strided_conv_as_two_tap_conv(stride):
    g = torch.Generator().manual_seed(10 + stride)
    c, c_out, L_in, B = 6, 7, 53, 2
    w = torch.randn(c_out, c, 2 * stride, generator=g)
    x = torch.randn(B, c, L_in, generator=g)
    pad = math.ceil(stride / 2)
    ref = F.conv1d(x, w, stride=stride, padding=pad)
    L_out = (L_in + 2 * pad - 2 * stride) // stride + 1
    assert ref.shape[-1] == L_out
    buf = torch.zeros(B, (L_out + 1) * stride, c)                               # `pad` zero rows in front, never-written tail stays zero
    n = min(L_in, (L_out + 1) * stride - pad)
    buf[:, pad:pad + n] = x.transpose(1, 2)[:, :n]
    view = buf.view(B, L_out + 1, stride * c)
    wp = w.permute(0, 2, 1).reshape(c_out, 2 * stride * c)                      # K index = tap * c + channel (DACEncoder packing)
    a = torch.cat([view[:, :-1], view[:, 1:]], dim=-1)                          # output t reads view rows t and t + 1
    torch.testing.assert_close((a @ wp.t()).transpose(1, 2), ref, rtol=1e-5, atol=1e-5)


def test_lengths_match_torch():
    from edm_tts_b200.dac_decoder import conv_transpose_length

    for s in (8, 5, 4, 2):
        for L_in in (1, 2, 17):
            assert conv_transpose_length(L_in, s) == F.conv_transpose1d(torch.zeros(1, 1, L_in), torch.zeros(1, 1, 2 * s), stride=s, padding=s // 2,
                                                                         output_padding=s % 2).shape[-1]


def test_dac_checkpoint_loader_reads_reference_layout(tmp_path):
    """DAC.load_checkpoint on an HF directory (config.json + model.safetensors) with the reference's key names."""
    import json

    from safetensors.torch import save_file

    from edm_tts_b200.dac import DAC
    from edm_tts_b200.synthetic import make_dac_state_dict

    sd = make_dac_state_dict(1)
    save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "model.safetensors"))
    (tmp_path / "config.json").write_text(json.dumps({"encoder_dim": 64, "encoder_rates": [2, 4, 5, 8], "decoder_dim": 1536, "decoder_rates": [8, 5, 4, 2],
                                                      "n_codebooks": 12, "codebook_size": 1024, "codebook_dim": 8, "sample_rate": 16000, "model_type": "dac"}))
    got, cfg = DAC.load_checkpoint(str(tmp_path))
    assert set(got) == set(sd) and cfg.latent_dim == 1024 and cfg.decoder_rates == (8, 5, 4, 2)
    k = "decoder.model.1.block.1.parametrizations.weight.original1"
    assert torch.equal(got[k], sd[k]) and got[k].shape == (1536, 768, 16)


def test_synthetic_dac_weights_follow_the_reference_key_layout(golden_dir):
    """The key names / shapes the CUDA-side DAC reads are those of the unmodified reference DAC(DACConfig()).state_dict()
    (tests/golden/dac_state_dict_keys.json, written by tests/golden/make_golden.py from /root/reference)."""
    import json
    import os

    from edm_tts_b200.synthetic import make_dac_state_dict

    with open(os.path.join(golden_dir, "dac_state_dict_keys.json")) as f:
        ref = json.load(f)
    ours = make_dac_state_dict(0)
    assert set(ours) == set(ref)
    assert all(list(ours[k].shape) == ref[k] for k in ref)


def test_code_lengths_match_a_conv_stack():
    """code_lengths == the length arithmetic of the reference encoder's Conv1d chain (what AudioTokenizer.get_code_lengths walks)."""
    from edm_tts_b200.dac_encoder import code_lengths

    convs = [torch.nn.Conv1d(1, 1, 7, padding=3)]
    for s in (2, 4, 5, 8):
        for d in (1, 3, 9):
            convs += [torch.nn.Conv1d(1, 1, 7, dilation=d, padding=3 * d), torch.nn.Conv1d(1, 1, 1)]
        convs.append(torch.nn.Conv1d(1, 1, 2 * s, stride=s, padding=math.ceil(s / 2)))
    convs.append(torch.nn.Conv1d(1, 1, 3, padding=1))
    net = torch.nn.Sequential(*convs)
    for n in (320, 321, 4000, 16000, 16077, 960160):
        with torch.inference_mode():
            assert net(torch.zeros(1, 1, n)).shape[-1] == code_lengths(n)
    lens = torch.tensor([320, 16077, 960160])
    assert code_lengths(lens).tolist() == [code_lengths(int(v)) for v in lens]
