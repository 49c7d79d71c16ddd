"""Generate golden vectors from the UNMODIFIED reference (build container only: needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference is imported from /root/reference (never copied). Harness-side patches, none of which touch reference files:
  * DAC.from_pretrained -> construct DAC(DACConfig(...)) (the reference's own from_pretrained breaks under transformers 5.x
    because DAC.__init__ never calls post_init());
  * torch.multinomial and Gumbel.sample read pre-generated noise so the run is bit-deterministic;
  * random_topk_mask / forward_first_level / encoder.forward are wrapped to record their outputs.
Weights are the deterministic ones of oracle/weights.py loaded into the reference modules, inputs oracle.weights.make_inputs.
Outputs: tests/golden/s2a_*.pt, rvq_*.pt, train_fwd_*.pt, dac_*.pt, t2s_*.pt (small tensors only).
"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True

from oracle.weights import OracleConfig, make_inputs, make_quantizer_state_dict, make_state_dict  # noqa: E402
from edm_tts_b200.synthetic import make_decoder_state_dict, make_encoder_state_dict  # noqa: E402

from edm_tts.models.dac import DAC  # noqa: E402
from edm_tts.models.dac.configuration import DACConfig  # noqa: E402
from edm_tts.models.dac.decoder import Decoder  # noqa: E402
from edm_tts.models.dac.encoder import Encoder  # noqa: E402
from edm_tts.models.dac.vector_quantizer import ResidualVectorQuantize  # noqa: E402
from edm_tts.models.injection_conformer import modeling_injection_conformer as mic  # noqa: E402
from edm_tts.models.injection_conformer.configuration import InjectionConformerConfig  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    # name: (OracleConfig, DACConfig kwargs)
    "small": (OracleConfig(hidden=128, heads=2, depth=6, injection_layers=(1, 2, 3, 4), num_semantic=50, n_codebooks=12,
                           codebook_size=64, codebook_dim=8, latent_dim=128),
              dict(encoder_dim=8, decoder_dim=32, codebook_size=64)),
    "full": (OracleConfig(), dict()),
}


def build_reference(cfg: OracleConfig, dac_kwargs, seed):
    dac_cfg = DACConfig(**dac_kwargs)
    DAC.from_pretrained = classmethod(lambda cls, path, *a, **k: cls(dac_cfg))  # shim, see module docstring
    rc = InjectionConformerConfig(hidden_size=cfg.hidden, num_semantic_tokens=cfg.num_semantic, acoustic_model_path="unused",
                                  encoder_num_heads=cfg.heads, encoder_num_layers=cfg.depth, encoder_ff_mult=cfg.ff_mult,
                                  encoder_conv_kernel_size=cfg.conv_kernel, injection_layers=list(cfg.injection_layers),
                                  residual=cfg.residual, use_injection=cfg.use_injection)
    model = mic.InjectionConformerModel(rc).eval()
    sd = make_state_dict(cfg, seed)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    bad = [k for k in res.missing_keys if not (k.startswith("acoustic_model.encoder.") or k.startswith("acoustic_model.decoder."))]
    assert not bad, f"hot-path keys not covered by oracle.weights: {bad[:5]}"
    return model, sd


class NoiseFeed:
    """Replaces the reference's two RNG consumers with reads from pre-generated tensors and records decisions."""

    def __init__(self, cat_gumbel, remask_gumbel):
        self.cat, self.rem = cat_gumbel, remask_gumbel
        self.i_cat = self.i_rem = 0
        self.ids, self.masks, self.first_logits = [], [], []

    def multinomial(self, probs, num_samples, replacement=False, *, generator=None):
        g = self.cat[self.i_cat]
        self.i_cat += 1
        return (torch.log(probs) + g).argmax(dim=-1, keepdim=True)

    def gumbel_sample(self, dist_self, sample_shape=torch.Size()):
        g = self.rem[self.i_rem]
        self.i_rem += 1
        return g.unsqueeze(-1)


def run_reference(model, inp, steps, temperature, autocast=False):
    feed = NoiseFeed(inp["cat_gumbel"], inp["remask_gumbel"])
    orig_multinomial, orig_gumbel = torch.multinomial, torch.distributions.gumbel.Gumbel.sample
    orig_topk, orig_ffl, orig_fwd = mic.random_topk_mask, model.encoder.forward_first_level, model.encoder.forward
    rec = {}

    def topk(*a, **k):
        m = orig_topk(*a, **k)
        feed.masks.append(m.clone())
        return m

    def ffl(*a, **k):
        out = orig_ffl(*a, **k)
        feed.first_logits.append(out[:, 0].float().clone())
        return out

    def fwd(*a, **k):
        out = orig_fwd(*a, **k)
        rec["all_logits"] = out.float().clone()
        return out

    torch.multinomial = feed.multinomial
    torch.distributions.gumbel.Gumbel.sample = feed.gumbel_sample
    mic.random_topk_mask = topk
    model.encoder.forward_first_level = ffl
    model.encoder.forward = fwd
    try:
        with torch.inference_mode(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            codes = model.infer_special(inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"],
                                        steps=steps, temperature=temperature)
    finally:
        torch.multinomial = orig_multinomial
        torch.distributions.gumbel.Gumbel.sample = orig_gumbel
        mic.random_topk_mask = orig_topk
        model.encoder.forward_first_level = orig_ffl
        model.encoder.forward = orig_fwd
    rec.update(codes=codes, step_masks=feed.masks, first_logits=feed.first_logits)
    return rec


def subsample_rows(t, n=6):
    """Keep n evenly spaced target rows of a [B, T, V] tensor."""
    idx = torch.linspace(0, t.shape[1] - 1, n).long()
    return t[:, idx].clone(), idx


def make_s2a(name, cfg_name, B, T, P, steps, seed=0, temperature=1.0, with_autocast=False):
    cfg, dac_kwargs = CONFIGS[cfg_name]
    torch.manual_seed(0)
    model, _ = build_reference(cfg, dac_kwargs, seed)
    inp = make_inputs(B, T, P, steps, cfg, seed=1234 + T + 7 * P + steps)
    rec = run_reference(model, inp, steps, temperature)
    gold = dict(cfg_name=cfg_name, B=B, T=T, P=P, steps=steps, weight_seed=seed, input_seed=1234 + T + 7 * P + steps, temperature=temperature,
                codes=rec["codes"].to(torch.int16), step_masks=[m.clone() for m in rec["step_masks"]],
                step_argmax=[l.argmax(-1).to(torch.int16) for l in rec["first_logits"]])
    rows = [subsample_rows(l) for l in rec["first_logits"]]
    gold["step_logit_rows"] = [r[0] for r in rows]
    gold["row_idx"] = rows[0][1] if rows else None
    al = rec["all_logits"]  # [B, Q, T, V]
    idx = torch.linspace(0, T - 1, 4).long()
    gold["final_logit_rows"] = al[:, :, idx].clone()
    gold["final_row_idx"] = idx
    top2 = al.topk(2, dim=-1)[0]
    gold["final_margin"] = (top2[..., 0] - top2[..., 1]).to(torch.float16)   # reference's own top-1 margin per decision
    if with_autocast:
        rec_bf = run_reference(model, inp, steps, temperature, autocast=True)
        gold["bf16_step0_logit_rows"] = rec_bf["first_logits"][0][:, rows[0][1]].clone() if rows else None
        gold["bf16_codes"] = rec_bf["codes"].to(torch.int16)
    torch.save(gold, os.path.join(OUT, f"s2a_{name}.pt"))
    print(f"s2a_{name}: codes {tuple(rec['codes'].shape)} steps={steps} saved", flush=True)


def make_rvq(name, cfg_name, B, T, seed=0):
    cfg, dac_kwargs = CONFIGS[cfg_name]
    q = ResidualVectorQuantize(input_dim=cfg.latent_dim, n_codebooks=cfg.n_codebooks, codebook_size=cfg.codebook_size,
                               codebook_dim=cfg.codebook_dim, quantizer_dropout=0.5).eval()
    q.load_state_dict(make_quantizer_state_dict(cfg, seed), strict=True)
    g = torch.Generator().manual_seed(99 + T)
    z = torch.randn(B, cfg.latent_dim, T, generator=g)
    with torch.inference_mode():
        out = q(z)
        feats = q.from_codes(out["codes"])[0]
        unred = q.from_codes_unreduced(out["codes"][:, :4])
    gold = dict(cfg_name=cfg_name, B=B, T=T, weight_seed=seed, z_seed=99 + T, codes=out["codes"].to(torch.int16),
                latents_head=out["latents"][:, :, :8].clone(), zq_head=out["z"][:, :16, :8].clone(),
                feats_head=feats[:, :16, :8].clone(), unred_head=unred[:, :, :16, :8].clone(),
                zq_sum=out["z"].double().sum().item(), feats_sum=feats.double().sum().item())
    torch.save(gold, os.path.join(OUT, f"rvq_{name}.pt"))
    print(f"rvq_{name}: codes {tuple(out['codes'].shape)} saved", flush=True)


def make_dac_encoder(name, encoder_dim, B, L, seed=0):
    """The reference Encoder (fp32, CPU) on synthetic audio: full z for small shapes, a slice + checksums otherwise."""
    rates = (2, 4, 5, 8)
    enc = Encoder(encoder_dim, list(rates)).eval()
    enc.load_state_dict(make_encoder_state_dict(encoder_dim, rates, seed), strict=True)
    audio = (torch.randn(B, 1, L, generator=torch.Generator().manual_seed(7 + L)) * 0.3).clamp(-1, 1)
    with torch.inference_mode():
        z = enc(audio)
    gold = dict(encoder_dim=encoder_dim, B=B, L=L, weight_seed=seed, audio_seed=7 + L, z_shape=tuple(z.shape),
                z=z.clone() if z.numel() <= 1 << 18 else None, z_head=z[:, :32, :16].clone(), z_tail=z[:, -32:, -16:].clone(),
                z_sum=z.double().sum().item(), z_abs_sum=z.double().abs().sum().item())
    torch.save(gold, os.path.join(OUT, f"dac_encoder_{name}.pt"))
    print(f"dac_encoder_{name}: z {tuple(z.shape)} rms {z.pow(2).mean().sqrt().item():.3f} saved", flush=True)


def make_dac_decoder(name, input_channel, channels, B, T, seed=0):
    """The reference Decoder (fp32, CPU) on a synthetic latent: audio head / tail / strided samples and checksums."""
    rates = (8, 5, 4, 2)
    dec = Decoder(input_channel, channels, list(rates)).eval()
    dec.load_state_dict(make_decoder_state_dict(input_channel, channels, rates, seed), strict=True)
    z = torch.randn(B, input_channel, T, generator=torch.Generator().manual_seed(11 + T)) * 0.5
    with torch.inference_mode():
        audio = dec(z)
    gold = dict(input_channel=input_channel, channels=channels, B=B, T=T, weight_seed=seed, z_seed=11 + T, audio_shape=tuple(audio.shape),
                audio=audio.clone() if audio.numel() <= 1 << 17 else None, audio_head=audio[:, :, :256].clone(), audio_tail=audio[:, :, -256:].clone(),
                audio_strided=audio[:, :, ::37].clone(), audio_sum=audio.double().sum().item(), audio_abs_sum=audio.double().abs().sum().item())
    torch.save(gold, os.path.join(OUT, f"dac_decoder_{name}.pt"))
    print(f"dac_decoder_{name}: audio {tuple(audio.shape)} rms {audio.pow(2).mean().sqrt().item():.3f} saved", flush=True)


def make_dac_key_layout():
    """Key names and shapes of the unmodified reference DAC(DACConfig()).state_dict(): the checkpoint layout the CUDA-side DAC reads."""
    import json

    ref = DAC(DACConfig()).state_dict()
    with open(os.path.join(OUT, "dac_state_dict_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in ref.items()}, f, indent=0, sort_keys=True)
    print(f"dac_state_dict_keys: {len(ref)} entries saved", flush=True)


def make_train_forward(name, cfg_name, B, T, seed=0, loss_all=False):
    """InjectionConformerModel.forward (eval mode: no dropout, ground-truth injections) with cosine_schedule_mask replaced by a
    fixed Bernoulli(0.6) mask: loss, arg-max codes and a few logit rows."""
    cfg, dac_kwargs = CONFIGS[cfg_name]
    torch.manual_seed(0)
    model, _ = build_reference(cfg, dac_kwargs, seed)
    model.loss_all = loss_all                    # config.loss_all (modeling_injection_conformer.py:60)
    g = torch.Generator().manual_seed(4321 + T)
    sem = torch.randint(0, cfg.num_semantic, (B, T), generator=g)
    ac = torch.randint(0, cfg.codebook_size, (B, cfg.n_codebooks, T), generator=g)
    mask = torch.rand(B, T, generator=g) < 0.6
    rec = {}
    orig_fwd = model.encoder.forward

    def fwd(*a, **k):
        out = orig_fwd(*a, **k)
        rec["all_logits"] = out.float().clone()
        return out

    model.encoder.forward = fwd
    model.cosine_schedule_mask = lambda feature_length, batch_size: mask
    with torch.inference_mode():
        out = model(ac, sem)
    idx = torch.linspace(0, T - 1, 4).long()
    al = rec["all_logits"]
    top2 = al.topk(2, dim=-1)[0]
    gold = dict(cfg_name=cfg_name, B=B, T=T, weight_seed=seed, semantic_tokens=sem.to(torch.int16), acoustic_tokens=ac.to(torch.int16), mask=mask,
                loss=out.loss.item(), output_codes=out.output_acoustic_codes.to(torch.int16), logit_rows=al[:, :, idx].clone(), row_idx=idx,
                margin=(top2[..., 0] - top2[..., 1]).to(torch.float16), loss_all=loss_all)
    torch.save(gold, os.path.join(OUT, f"train_fwd_{name}.pt"))
    print(f"train_fwd_{name}: loss {out.loss.item():.6f} codes {tuple(out.output_acoustic_codes.shape)} saved", flush=True)


from tests.golden.make_golden_cfg import T2S_CONFIGS  # noqa: E402


def make_t2s(name, cfg_name, text, pred_iters, gt_length=None, seed=0):
    """TextToSemanticWLen.infer of the unmodified reference with injected sampling noise: predicted length, per-iteration own ids /
    masks, the final tokens, and logit rows of the first and last iteration."""
    from edm_tts.models.text_to_semantic.configuration import TextToSemanticWLenConfig
    from edm_tts.models.text_to_semantic import modeling_text_to_semantic as mts
    from edm_tts_b200.synthetic import T2SConfig, make_t2s_noise, make_t2s_state_dict

    cfg = T2SConfig(**T2S_CONFIGS[cfg_name])
    rc = TextToSemanticWLenConfig(hidden_size=cfg.hidden, semantic_vocab_size=cfg.semantic_vocab, text_vocab_size=cfg.text_vocab,
                                  main_encoder_num_heads=cfg.heads, main_encoder_num_layers=cfg.depth, length_predictor_num_heads=cfg.lp_heads,
                                  length_predictor_num_layers=cfg.lp_depth)
    torch.manual_seed(0)
    model = mts.TextToSemanticWLen(rc).eval()
    sd = make_t2s_state_dict(cfg, seed)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    buffers = {"text_token", "speech_token", "sep_token", "pad_token", "mask_token", "false"}
    bad = [k for k in res.missing_keys if k not in buffers and not k.endswith("rotary_emb.inv_freq")]
    assert not bad, f"keys not covered by make_t2s_state_dict: {bad[:5]}"
    # the length first (its own forward), then the noise for that sequence length
    n_text = len(text.encode("utf-8"))
    rec = {"logits": []}
    with torch.inference_mode():
        if gt_length is None:
            tt = torch.tensor(list(text.encode("utf-8")), dtype=torch.long) + model.num_special_tokens
            lp_in = torch.cat([model.length_token, model.input_embedding(tt).unsqueeze(0)], dim=1)
            raw = model.length_pred_head(model.length_predictor(lp_in, return_attn=False)[0][:, 0]).squeeze(-1)
            length = int(raw.exp().ceil().long().item())
        else:
            raw, length = None, int(gt_length)
    L = n_text + length + 4
    noise = make_t2s_noise(L, pred_iters, cfg, seed=777 + L)
    feed = NoiseFeed(noise["cat_gumbel"], noise["remask_gumbel"][:, 0])
    orig_mult, orig_gs = torch.multinomial, torch.distributions.gumbel.Gumbel.sample
    orig_e2l = model.embeddings_to_logits
    orig_topk = mts.random_topk_mask
    masks = []

    def e2l(*a, **k):
        out = orig_e2l(*a, **k)
        rec["logits"].append(out.float().clone())
        return out

    def topk(*a, **k):
        m = orig_topk(*a, **k)
        masks.append(m.clone())
        return m

    def gumbel_sample(dist_self, sample_shape=torch.Size()):
        g = feed.rem[feed.i_rem]
        feed.i_rem += 1
        return g.view(*sample_shape, 1)

    model.embeddings_to_logits = e2l
    mts.random_topk_mask = topk
    torch.multinomial = feed.multinomial
    torch.distributions.gumbel.Gumbel.sample = gumbel_sample
    try:
        with torch.inference_mode():
            out = model.infer(text=text, pred_iters=pred_iters, temperature=1.0, gt_length=gt_length)
    finally:
        torch.multinomial, torch.distributions.gumbel.Gumbel.sample = orig_mult, orig_gs
        mts.random_topk_mask = orig_topk
    assert feed.i_cat == pred_iters - 1 and feed.i_rem == pred_iters - 1
    tokens = out.speech_pred_tokens
    first, last = rec["logits"][0][0], rec["logits"][-1][0]
    rows = torch.linspace(0, L - 1, 6).long()
    top2 = last.topk(2, dim=-1)[0]
    gold = dict(cfg_name=cfg_name, text=text, pred_iters=pred_iters, gt_length=gt_length, weight_seed=seed, length=length,
                raw_log_length=None if raw is None else raw.item(), noise_seed=777 + L, tokens=tokens.to(torch.int16),
                masks=torch.stack(masks) if masks else None, row_idx=rows, first_rows=first[rows].clone(), last_rows=last[rows].clone(),
                last_margin=(top2[:, 0] - top2[:, 1]).to(torch.float16))
    torch.save(gold, os.path.join(OUT, f"t2s_{name}.pt"))
    print(f"t2s_{name}: length {length} tokens {tuple(tokens.shape)} masks {[int(m.sum()) for m in masks]} saved", flush=True)


def make_all_t2s():
    make_t2s("small_s4", "small", "hello world", 4)
    make_t2s("small_s1", "small", "one step", 1, gt_length=23)
    make_t2s("base_s3", "base", "The quick brown fox.", 3)
    make_t2s("train_s4", "train", "Grafted onto new rootstock, the old tree bore fruit.", 4, gt_length=75)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "t2s":
        make_all_t2s()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dac_encoder":      # only the encoder fixtures (added after the others)
        make_dac_encoder("small", 8, 2, 3200 + 137)
        make_dac_encoder("full", 64, 2, 6400 + 160)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "loss_all":          # the loss_all branch of forward (returns [b, q, t] codes)
        make_train_forward("small_loss_all", "small", 2, 33, loss_all=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dac_decoder":
        make_dac_decoder("small", 64, 96, 2, 9)
        make_dac_decoder("full", 1024, 1536, 2, 12)
        make_dac_key_layout()
        sys.exit(0)
    with tempfile.TemporaryDirectory():
        make_rvq("small", "small", 2, 50)
        make_rvq("full", "full", 2, 75)
        make_s2a("small_s1", "small", 2, 40, 0, 1)
        make_s2a("small_s4", "small", 2, 40, 0, 4)
        make_s2a("small_s8_prompt", "small", 2, 40, 16, 8, with_autocast=True)
        make_s2a("full_s1", "full", 1, 60, 0, 1)
        make_s2a("full_s8", "full", 2, 150, 0, 8, with_autocast=True)
        make_s2a("full_s4_prompt", "full", 1, 100, 50, 4)
        make_train_forward("small", "small", 2, 40)
        make_train_forward("full", "full", 1, 60)
        make_train_forward("small_loss_all", "small", 2, 33, loss_all=True)
        make_dac_encoder("small", 8, 2, 3200 + 137)
        make_dac_encoder("full", 64, 2, 6400 + 160)
        make_dac_decoder("small", 64, 96, 2, 9)
        make_dac_decoder("full", 1024, 1536, 2, 12)
        make_dac_key_layout()
        make_all_t2s()
