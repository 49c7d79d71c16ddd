"""GPU parity of the S2A decode path (through the C ABI) against the oracle and the reference's golden vectors.

Tolerances (bf16 tensor-core GEMMs with fp32 accumulation vs the fp32 oracle; logits are ~unit variance):
  first-level / final logits: max |diff| < 0.15, mean |diff| < 0.02;
  every discrete disagreement -- arg-max of the logits, the path's own sampled id under the same injected Gumbel noise -- must be a
  near-tie: the oracle's own margin between its choice and ours is < NEAR_TIE_EPS = 0.08 (tests/parity_utils.py derives it from the
  measured logit error; flips above SURVEY.md's suggested 0.06 are counted separately), teacher-forced upstream; the path's own
  re-masking decision may differ only for tokens within MASK_TIE_EPS = 0.06 of the oracle's cut-off confidence. Exceptions are
  counted and listed per stage (profiles/parity_r02.json).
Index / mask work that does not depend on bf16 logits (encoder input, code assembly) is exact.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

MAX_TOL, MEAN_TOL = 0.15, 0.02


def _check_reports(reports):
    for r in reports:
        head = f"{r['what']:32s} " + (f"max={r['max']:.4f} mean={r['mean']:.5f} " if "max" in r else " " * 24)
        print(head + f"agree={r['agree']:.5f} mismatch={r['n_mismatch']}/{r['n']} not_near_tie={r['n_not_near_tie']} "
              f"worst_margin={r['worst_margin']:.4f}")
    for r in reports:
        if "max" in r:
            assert r["max"] < MAX_TOL and r["mean"] < MEAN_TOL, r
        assert r["n_not_near_tie"] == 0, r
        assert r["agree"] > 0.9 or r["n"] < 30, r


@pytest.mark.parametrize("B,T,P,steps", [(2, 150, 0, 8), (1, 60, 0, 1), (1, 100, 50, 4), (3, 77, 33, 2),
                                         (1, 1, 0, 1), (2, 3, 2, 2), (1, 600, 0, 3), (1, 1500, 150, 2)])
def test_teacher_forced_parity_vs_oracle(B, T, P, steps):
    from tests.parity_utils import full_model, teacher_forced_parity

    cfg, sd, model = full_model()
    inp, ref, ours, reports = teacher_forced_parity(cfg, sd, model, B, T, P, steps, input_seed=1234 + T + 7 * P + steps)
    # encoder input: gathers + LayerNorm only -> tight
    torch.testing.assert_close(ours["x0"], ref["x0"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ours["x_final"], ref["x_final"], rtol=1e-4, atol=2e-4)
    _check_reports(reports)


@pytest.mark.parametrize("B,T,P,steps", [(64, 500, 0, 8), (8, 1500, 150, 2)])
def test_teacher_forced_parity_at_bench_size(B, T, P, steps):
    """The same protocol at the sizes the bench runs (BASELINE config 2: 64 x 500 frames x 8 steps; config 5: 8 x 30 s with a 3 s
    prompt): every GEMM is on the CTA-pair kernel here (TMA-reduce residual epilogue, TMA-store logits / Swish / RoPE epilogues),
    attention and the conv module run their multi-tile paths. The oracle runs in fp32 on the same GPU."""
    from tests.parity_utils import full_model, teacher_forced_parity

    cfg, sd, model = full_model()
    inp, ref, ours, reports = teacher_forced_parity(cfg, sd, model, B, T, P, steps, input_seed=99 + B + T)
    torch.testing.assert_close(ours["x0"], ref["x0"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ours["x_final"], ref["x_final"], rtol=1e-4, atol=2e-4)
    _check_reports(reports)


def test_chunked_decode_parity_vs_oracle():
    """B = 130 > MAX_CHUNK: three workspaces' worth (64 + 64 + 2, ragged last chunk) through infer_special itself, teacher-forced
    with the oracle's decisions; the returned codes must equal the oracle's except where the oracle's final-logit margin is a near-tie."""
    from oracle.weights import make_inputs
    from tests.parity_utils import NEAR_TIE_EPS, compare_ids, full_model, oracle_trace, record

    cfg, sd, model = full_model()
    B, T, steps = 130, 60, 3
    inp = make_inputs(B, T, 0, steps, cfg, seed=4242)
    ref = oracle_trace(cfg, sd, inp, steps)
    codes = model.infer_special(inp["semantic_tokens"], None, None, steps=steps, cat_gumbel=inp["cat_gumbel"], remask_gumbel=inp["remask_gumbel"],
                                forced_ids=torch.stack(ref["step_ids"]), forced_masks=torch.stack(ref["step_masks"]),
                                forced_coarse=ref["all_logits"][:, :4].argmax(-1))
    r = compare_ids(codes, ref["codes"], ref["all_logits"], "codes of a 3-chunk decode (B=130)", NEAR_TIE_EPS)
    record("B=130 T=60 P=0 S=3 (chunked infer_special)", **r)
    print(r)
    assert r["n_not_near_tie"] == 0 and r["agree"] > 0.97, r
    # the ragged last chunk alone, with its global batch offset, reproduces its rows
    tail = model.infer_special(inp["semantic_tokens"][128:], None, None, steps=steps, batch_offset=128,
                               cat_gumbel=inp["cat_gumbel"].view(steps - 1, B, T, -1)[:, 128:].reshape(steps - 1, 2 * T, -1),
                               remask_gumbel=inp["remask_gumbel"][:, 128:], forced_ids=torch.stack(ref["step_ids"])[:, 128:],
                               forced_masks=torch.stack(ref["step_masks"])[:, 128:], forced_coarse=ref["all_logits"][128:, :4].argmax(-1))
    assert torch.equal(tail, codes[128:])


def test_encoder_methods_vs_oracle():
    """The mirror's encoder API called the way the reference model calls it: forward_first_level / forward (with feature-valued
    `injections`, as infer_special builds them, and with prompt codes) / apply_single_to_logits, plus model.semantic_embedding,
    model.acoustic_feat_proj and model.acoustic_model.codes_to_features*."""
    from oracle import s2a as os2a
    from oracle.weights import make_inputs
    from tests.parity_utils import compare_logits, full_model, record

    cfg, sd, model = full_model()
    B, T, P = 2, 90, 30
    inp = make_inputs(B, T, P, 1, cfg, seed=31)
    sem, ap, sp = (inp[k].cuda() for k in ("semantic_tokens", "acoustic_prompt_tokens", "semantic_prompt_tokens"))
    with torch.inference_mode():
        x, _, prompt_inj, _ = os2a.build_encoder_input(sd, cfg, sem, ap, sp)
        ref_first = os2a.forward_first_level(sd, cfg, x.clone(), P)
        ref_all = os2a.forward_full(sd, cfg, x.clone(), prompt_inj, P)
    # the same encoder input assembled from the mirror's own attributes (modeling_injection_conformer.py:139-168)
    feats = model.acoustic_model.codes_to_features_unreduced(ap)                                   # [B, q, 1024, P]
    torch.testing.assert_close(feats, os2a.codes_to_features_unreduced(sd, cfg, ap), rtol=1e-4, atol=1e-4)
    ac = model.acoustic_feat_proj(feats[:, 0].transpose(1, 2))
    x_ours = torch.cat([model.semantic_embedding(sp) + ac, model.semantic_embedding(sem) + model.mask_token], dim=1)
    assert (x_ours - x).abs().max().item() < 0.05                                                   # bf16 GEMM vs fp32 Linear before a LayerNorm
    mti = torch.zeros(B, P + T, dtype=torch.bool, device="cuda")
    mti[:, P:] = True
    first = model.encoder.forward_first_level(x, mask_time_indices=mti)
    assert first.shape == (B, 1, T, cfg.codebook_size)
    r = compare_logits(first[:, 0], ref_first, "encoder.forward_first_level")
    record("encoder API B=2 T=90 P=30", **r)
    assert r["max"] < MAX_TOL and r["n_not_near_tie"] == 0, r
    forced = ref_all[:, :4].argmax(-1)
    injections = [torch.cat([u, torch.zeros(B, T, 1024, device="cuda")], dim=1) for u in
                  (os2a.codes_to_features(sd, cfg, ap[:, : k + 1]).transpose(1, 2) for k in range(4))]
    for name, kw in (("injections", dict(injections=injections)), ("prompt_codes", dict(prompt_codes=ap))):
        logits = model.encoder(x, injections=kw.get("injections"), acoustic_model=model.acoustic_model, mask_time_indices=mti,
                               prompt_codes=kw.get("prompt_codes"), forced_coarse=forced)
        assert logits.shape == (B, cfg.n_codebooks, T, cfg.codebook_size)
        for q in range(cfg.n_codebooks):
            r = compare_logits(logits[:, q], ref_all[:, q], f"encoder.forward({name}) level {q}")
            record("encoder API B=2 T=90 P=30", **r)
            assert r["max"] < MAX_TOL and r["mean"] < MEAN_TOL and r["n_not_near_tie"] == 0, r
    with pytest.raises(ValueError):
        model.encoder(x, mask_time_indices=mti)                                                      # prompt rows but no injections
    # apply_single_to_logits on an arbitrary activation
    h = torch.randn(2, 40, 1024, device="cuda")
    for idx in (0, 5, 11):
        got = model.encoder.apply_single_to_logits(h, idx)
        want = os2a.single_to_logits(sd, h, idx, "fp32")
        assert got.shape == (2, 1, 40, cfg.codebook_size)
        assert (got[:, 0] - want).abs().max().item() < 0.05
    with pytest.raises(IndexError):
        model.encoder.apply_single_to_logits(h, 12)


def test_from_pretrained_roundtrip(tmp_path):
    """HF directory (config.json + model.safetensors, the reference's save_pretrained layout) -> InjectionConformerModel.from_pretrained;
    the loaded model decodes bit-identically to one built from the in-memory state dict."""
    import json

    from safetensors.torch import save_file

    from edm_tts_b200 import InjectionConformerModel
    from edm_tts_b200.config import InjectionConformerConfig
    from oracle.weights import OracleConfig, make_inputs, make_state_dict

    cfg = OracleConfig(depth=6, injection_layers=(1, 2, 3, 4))
    sd = make_state_dict(cfg, seed=3)
    dac_dir, s2a_dir = tmp_path / "dac", tmp_path / "s2a"
    dac_dir.mkdir()
    s2a_dir.mkdir()
    (dac_dir / "config.json").write_text(json.dumps(dict(encoder_dim=64, encoder_rates=[2, 4, 5, 8], n_codebooks=12, codebook_size=1024, codebook_dim=8)))
    (s2a_dir / "config.json").write_text(json.dumps(dict(
        hidden_size=1024, num_semantic_tokens=1024, acoustic_model_path=str(dac_dir), injection_layers=[1, 2, 3, 4], residual=True,
        use_injection=True, loss_all=False, encoder_config=dict(depth=6, heads=16, ff_mult=4, conv_kernel_size=5, dim_head=64))))
    save_file({k: v.contiguous() for k, v in sd.items()}, str(s2a_dir / "model.safetensors"))
    loaded = InjectionConformerModel.from_pretrained(str(s2a_dir)).eval().to("cuda")
    direct = InjectionConformerModel(InjectionConformerConfig(encoder_config=dict(depth=6, heads=16, ff_mult=4, conv_kernel_size=5),
                                                              injection_layers=(1, 2, 3, 4)), sd)
    assert loaded.injection_layers == [1, 2, 3, 4] and loaded.num_quantizers == 12
    inp = make_inputs(2, 40, 10, 3, cfg, seed=8)
    args = (inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"])
    assert torch.equal(loaded.infer_special(*args, steps=3, seed=1), direct.infer_special(*args, steps=3, seed=1))
    assert loaded.generate is not None and loaded.training is False
    with pytest.raises(NotImplementedError):
        loaded.train()
    with pytest.raises(ValueError):
        loaded.to("cpu")
    bad = {k: v for k, v in sd.items() if k != "encoder.fine_head.0.weight"}
    with pytest.raises(KeyError):
        InjectionConformerModel(direct.config, bad)                                                 # strict, like load_state_dict(strict=True)


def test_index_validation_and_device_guard():
    """Out-of-range tokens raise IndexError (F.embedding does in the reference) instead of faulting on the device; a model keeps working
    when another device / stream is current."""
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    inp = make_inputs(1, 20, 5, 2, cfg, seed=2)
    sem, ap, sp = inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"]
    bad = sem.clone()
    bad[0, 3] = cfg.num_semantic
    with pytest.raises(IndexError):
        model.infer_special(bad, None, None)
    with pytest.raises(IndexError):
        model.infer_special(bad.cuda(), None, None)
    badp = ap.clone()
    badp[0, 0, 0] = -1
    with pytest.raises(IndexError):
        model.infer_special(sem, badp, sp)
    with pytest.raises(IndexError):
        model.acoustic_model.codes_to_features(torch.full((1, 2, 4), 1024))
    with pytest.raises(IndexError):
        model.infer_special(sem[:, :1], None, None, steps=4)                                        # T = 1 cannot be re-masked
    ref = model.infer_special(sem, ap, sp, steps=2, seed=3)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        again = model.infer_special(sem, ap, sp, steps=2, seed=3)
    torch.cuda.synchronize()
    assert torch.equal(ref, again)
    if torch.cuda.device_count() > 1:
        with torch.cuda.device(1):                                                                   # model lives on cuda:0
            other = model.infer_special(sem, ap, sp, steps=2, seed=3)
        assert other.device == ref.device and torch.equal(other, ref)


@pytest.mark.parametrize("name", ["full_s1", "full_s8", "full_s4_prompt"])
def test_against_reference_golden(name, golden_dir):
    """Free-running CUDA decode vs the codes the unmodified reference produced (fp32, CPU). Without teacher forcing a
    single near-tie flip cascades, so the bar is per-level agreement, not equality; the first level of a 1-step decode has no
    cascade and must agree except at near-ties (reference margin from the golden file)."""
    from oracle.weights import make_inputs
    from tests.parity_utils import NEAR_TIE_EPS, full_model

    g = torch.load(os.path.join(golden_dir, f"s2a_{name}.pt"))
    cfg, sd, model = full_model(g["weight_seed"])
    inp = make_inputs(g["B"], g["T"], g["P"], g["steps"], cfg, seed=g["input_seed"])
    codes = model.infer_special(inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"], steps=g["steps"],
                                temperature=g["temperature"], cat_gumbel=inp["cat_gumbel"] if g["steps"] > 1 else None,
                                remask_gumbel=inp["remask_gumbel"] if g["steps"] > 1 else None)
    ref = g["codes"].long().to(codes.device)
    assert codes.shape == ref.shape and codes.dtype == torch.int64
    agree = (codes == ref).float().mean(dim=(0, 2))
    print(name, "per-level agreement with the reference:", [round(a, 3) for a in agree.tolist()])
    if g["steps"] == 1:
        mism = codes[:, 0] != ref[:, 0]
        margin = g["final_margin"][:, 0].float().to(codes.device)
        assert (margin[mism] < NEAR_TIE_EPS).all(), margin[mism]
        assert agree[0] > 0.93
    # multi-step free-running decodes diverge chaotically after the first flipped sample (random-init network); the strict
    # multi-step check is the teacher-forced test above, this one only guards against gross errors
    assert agree.mean() > (0.5 if g["steps"] == 1 else 0.15)


def test_batch_chunking_and_determinism(monkeypatch):
    """B > MAX_CHUNK is decoded in chunks; rows must not depend on which chunk / batch they were decoded in (this is
    what makes N-GPU batch sharding bit-identical to 1 GPU)."""
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    inp = make_inputs(5, 64, 0, 3, cfg, seed=7)
    kw = dict(steps=3, cat_gumbel=inp["cat_gumbel"], remask_gumbel=inp["remask_gumbel"])
    all_codes = model.infer_special(inp["semantic_tokens"], None, None, **kw)
    again = model.infer_special(inp["semantic_tokens"], None, None, **kw)
    assert torch.equal(all_codes, again)
    sub = slice(2, 4)
    V = cfg.codebook_size
    part = model.infer_special(inp["semantic_tokens"][sub], None, None, steps=3,
                               cat_gumbel=inp["cat_gumbel"].view(2, 5, 64, V)[:, sub].reshape(2, -1, V),
                               remask_gumbel=inp["remask_gumbel"][:, sub])
    assert torch.equal(part, all_codes[sub])
    import edm_tts_b200.s2a as s2a_mod

    monkeypatch.setattr(s2a_mod, "MAX_CHUNK", 2)
    chunked = model.infer_special(inp["semantic_tokens"], None, None, **kw)
    assert torch.equal(chunked, all_codes)


def test_fused_argmax_heads_equal_materialised_logits():
    """infer_special never materialises the [b, 12, t, 1024] logits (arg-max in the head GEMM epilogues, 16 partials per row); the
    codes must equal, bit for bit, those of the staged run that writes fp32 logits and arg-maxes them with sample_kernel. Checked on
    the small-M GEMM kernel (B=2) and on the CTA-pair kernel (B=16 x 200 frames), with a prompt and teacher-forced coarse tokens."""
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    for B, T, P, steps in ((2, 70, 10, 3), (16, 200, 0, 2)):
        inp = make_inputs(B, T, P, steps, cfg, seed=B + T)
        args = (inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"])
        tr = model.decode_trace(*args, steps=steps, seed=3)
        assert torch.equal(tr["codes"], tr["all_logits"].argmax(-1))
        fused = model.infer_special(*args, steps=steps, seed=3)
        assert torch.equal(fused, tr["codes"])
        fc = torch.randint(0, 1024, (B, 4, T))
        tr2 = model.decode_trace(*args, steps=steps, seed=3, forced_coarse=fc)
        assert torch.equal(model.infer_special(*args, steps=steps, seed=3, forced_coarse=fc), tr2["codes"])


def test_philox_path_runs_and_is_seeded():
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    inp = make_inputs(2, 50, 0, 1, cfg, seed=3)
    a = model.infer_special(inp["semantic_tokens"], None, None, steps=4, seed=11)
    b = model.infer_special(inp["semantic_tokens"], None, None, steps=4, seed=11)
    c = model.infer_special(inp["semantic_tokens"], None, None, steps=4, seed=12)
    assert torch.equal(a, b)
    assert not torch.equal(a, c)
    assert a.min() >= 0 and a.max() < cfg.codebook_size


def test_api_errors():
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    sem = torch.zeros(1, 10, dtype=torch.long)
    with pytest.raises(ValueError):
        model.infer_special(sem, torch.zeros(1, 12, 5, dtype=torch.long), torch.zeros(1, 6, dtype=torch.long))
    with pytest.raises(IndexError):
        model.infer_special(sem, torch.zeros(1, 2, 5, dtype=torch.long), torch.zeros(1, 5, dtype=torch.long))
    with pytest.raises(AssertionError):
        model.forward(torch.zeros(1, 12, 9, dtype=torch.long), sem)


@pytest.mark.parametrize("B,T,loss_all", [(2, 150, False), (3, 77, True), (1, 1, False)])
def test_eval_forward_loss_vs_oracle(B, T, loss_all):
    """InjectionConformerModel.forward (eval mode): loss and arg-max codes against the oracle restatement with the same mask."""
    from oracle import s2a as os2a
    from tests.parity_utils import compare_logits, full_model

    cfg, sd, model = full_model()
    g = torch.Generator().manual_seed(77 + T)
    sem = torch.randint(0, cfg.num_semantic, (B, T), generator=g)
    ac = torch.randint(0, cfg.codebook_size, (B, cfg.n_codebooks, T), generator=g)
    mask = torch.rand(B, T, generator=g) < 0.6
    mask[:, 0] = True                                            # keep at least one masked position per sequence
    with torch.inference_mode():
        ref = os2a.training_forward(sd, cfg, ac.cuda(), sem.cuda(), mask.cuda(), loss_all=loss_all)
    old = model.loss_all
    model.loss_all = loss_all
    try:
        out = model(ac, sem, mask_time_indices=mask)
    finally:
        model.loss_all = old
    torch.cuda.synchronize()
    print(f"loss ours {out.loss.item():.5f} oracle {ref['loss'].item():.5f}")
    assert abs(out.loss.item() - ref["loss"].item()) < 2e-2
    assert torch.equal(out["target_acoustic_codes"].cpu(), ac)
    assert out.output_acoustic_codes.shape == ref["output_acoustic_codes"].shape
    # arg-max codes agree except at near-ties of the oracle's own logits
    sel = (torch.ones_like(mask) if loss_all else mask).cuda()[:, None, :].expand(B, cfg.n_codebooks, T)
    ref_logits = ref["logits"].masked_select(sel[..., None]).view(-1, cfg.codebook_size)
    if loss_all:
        assert out.output_acoustic_codes.shape == (B, cfg.n_codebooks, T)      # un-flattened, as the reference returns it
    ours_codes, ref_codes = out.output_acoustic_codes.reshape(-1), ref["output_acoustic_codes"].reshape(-1)
    mism = ours_codes != ref_codes
    margin = ref_logits.gather(1, ref_codes[:, None])[:, 0] - ref_logits.gather(1, ours_codes[:, None])[:, 0]
    assert (margin[mism] < 0.08).all(), margin[mism].max().item()
    assert mism.float().mean().item() < 0.1 or mism.numel() < 30
    with pytest.raises(AssertionError):
        model(ac[..., :-1], sem) if T > 1 else model(ac, sem[..., :0])


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (B=64, T=500, 8 steps), where the oracle is too slow to run: size-independent properties.
    (1) determinism: same seed -> same codes; (2) shard invariance: rows decoded alone (with their global batch offset) equal
    the rows of the full batch; (3) codes are valid indices; (4) sequences with identical semantic tokens but different rows
    draw different noise (Philox counters are per global row), while a duplicated *row index* reproduces."""
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    sem = make_inputs(64, 500, 0, 1, cfg, seed=2024)["semantic_tokens"]
    full = model.infer_special(sem, None, None, steps=8, seed=5)
    again = model.infer_special(sem, None, None, steps=8, seed=5)
    assert torch.equal(full, again)
    assert full.shape == (64, 12, 500) and full.min() >= 0 and full.max() < cfg.codebook_size
    part = model.infer_special(sem[40:48], None, None, steps=8, seed=5, batch_offset=40)
    assert torch.equal(part, full[40:48])
    twin = sem.clone()
    twin[1] = twin[0]
    out = model.infer_special(twin[:2], None, None, steps=8, seed=5)
    assert not torch.equal(out[0], out[1])          # same tokens, different Philox rows
    assert torch.equal(out[0], full[0])             # row 0 unchanged by what sits next to it


def test_decode_is_cuda_graph_capturable():
    """infer_special issues no host synchronisation and no allocation outside torch's pool, so the whole decode (854 launches at the
    base config) can be captured with torch.cuda.graph and replayed; the replay reproduces the stream launch bit for bit."""
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    from oracle.weights import make_inputs

    inp = make_inputs(1, 60, 20, 3, cfg, seed=77)
    sem, ap, sp = inp["semantic_tokens"].cuda(), inp["acoustic_prompt_tokens"].cuda(), inp["semantic_prompt_tokens"].cuda()
    ref = model.infer_special(sem, ap, sp, steps=3, seed=5)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):          # the mirror orders this call after `ref` (one workspace per context), although the stream differs
        warm = model.infer_special(sem, ap, sp, steps=3, seed=5)
    torch.cuda.synchronize()
    if not torch.equal(warm, ref):  # diagnostics: which of the two is the odd one out, and where
        r2 = model.infer_special(sem, ap, sp, steps=3, seed=5)
        with torch.cuda.stream(s):
            w2 = model.infer_special(sem, ap, sp, steps=3, seed=5)
        torch.cuda.synchronize()
        frac = lambda a, b: [round(float(x), 3) for x in (a != b).float().mean(dim=(0, 2)).tolist()]
        raise AssertionError(f"side-stream decode differs: warm vs ref {frac(warm, ref)}; ref2 vs ref {frac(r2, ref)}; w2 vs warm {frac(w2, warm)}; "
                             f"w2 vs ref {frac(w2, ref)}; low_latency={model.low_latency}")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out = model.infer_special(sem, ap, sp, steps=3, seed=5)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    # new tokens through the captured graph: copy into the static input, replay
    inp2 = make_inputs(1, 60, 20, 3, cfg, seed=78)
    ref2 = model.infer_special(inp2["semantic_tokens"].cuda(), ap, sp, steps=3, seed=5)
    sem.copy_(inp2["semantic_tokens"].cuda())
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref2)


@pytest.mark.parametrize("low_latency", [None, False])
def test_graphed_decode_matches_infer_special(low_latency):
    """edm_tts_b200.serving.GraphedDecode: replays reproduce infer_special on the same tokens and the same (per-request) noise, in the
    mode the graph was captured in (low_latency=None: the single-utterance mode is chosen for these <= 512-row shapes)."""
    from edm_tts_b200.serving import GraphedDecode
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    gd = GraphedDecode(model, 2, 40, prompt_frames=10, steps=3, seed=9, low_latency=low_latency)
    assert gd.low_latency == (low_latency is None) and model.low_latency is False    # the capture leaves the model in its own mode
    model.set_low_latency(gd.low_latency)
    try:
        prev = None
        for s in (1, 2):
            inp = make_inputs(2, 40, 10, 3, cfg, seed=100 + s)
            got = gd(inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"])
            want = model.infer_special(inp["semantic_tokens"], inp["acoustic_prompt_tokens"], inp["semantic_prompt_tokens"], steps=3,
                                       cat_gumbel=gd.cat.clone(), remask_gumbel=gd.rem.clone())
            assert torch.equal(got, want) and got.shape == (2, 12, 40)
            assert prev is None or not torch.equal(prev, gd.cat)        # fresh noise per request
            prev = gd.cat.clone()
        with pytest.raises(ValueError):
            gd(torch.zeros(2, 41, dtype=torch.long))
        # Philox mode (no injected noise): request n equals infer_special with seed + n -- replays do not share noise
        gp = GraphedDecode(model, 1, 30, steps=2, seed=4, fresh_noise=False, low_latency=low_latency)
        tok = make_inputs(1, 30, 0, 2, cfg, seed=5)["semantic_tokens"]
        r0, r1 = gp(tok), gp(tok)
        assert torch.equal(r0, model.infer_special(tok, None, None, steps=2, seed=4))
        assert torch.equal(r1, model.infer_special(tok, None, None, steps=2, seed=5))
        assert not torch.equal(r0, r1)
        assert torch.equal(gp(tok, request_id=0), r0)
    finally:
        model.set_low_latency(False)


def test_calls_on_alternating_streams_are_ordered():
    """One workspace per context: a decode issued on another stream than the previous one must wait for it (the mirror records an event
    per call). Thirty alternations of the default stream and a fresh side stream reproduce the first result every time."""
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model

    cfg, sd, model = full_model()
    inp = make_inputs(1, 60, 20, 3, cfg, seed=77)
    sem, ap, sp = inp["semantic_tokens"].cuda(), inp["acoustic_prompt_tokens"].cuda(), inp["semantic_prompt_tokens"].cuda()
    first = model.infer_special(sem, ap, sp, steps=3, seed=5).clone()
    for i in range(30):
        s = torch.cuda.Stream()
        ref = model.infer_special(sem, ap, sp, steps=3, seed=5)
        with torch.cuda.stream(s):
            side = model.infer_special(sem, ap, sp, steps=3, seed=5)
        torch.cuda.synchronize()
        assert torch.equal(ref, first) and torch.equal(side, first), f"round {i}"


def test_low_latency_mode_parity_and_default_invariance():
    """set_low_latency(True) (edm_s2a_set_low_latency: split-K residual GEMMs for decodes of <= 512 rows, the reference's single-utterance
    call shape, inference.py:43-48): teacher-forced parity against the oracle under the same bars as the default mode, deterministic,
    and switched off again the model reproduces the default mode's codes bit for bit (whose rows never depend on their batch)."""
    from oracle.weights import make_inputs
    from tests.parity_utils import full_model, teacher_forced_parity

    cfg, sd, model = full_model()
    free = make_inputs(1, 150, 0, 8, cfg, seed=77)
    kw = dict(steps=8, cat_gumbel=free["cat_gumbel"], remask_gumbel=free["remask_gumbel"])
    base = model.infer_special(free["semantic_tokens"], None, None, **kw)
    model.set_low_latency(True)
    try:
        for (B, T, P, steps) in [(1, 150, 0, 8), (1, 100, 50, 4), (2, 150, 0, 2), (1, 500, 0, 2)]:
            inp, ref, ours, reports = teacher_forced_parity(cfg, sd, model, B, T, P, steps, input_seed=31 + T + P)
            torch.testing.assert_close(ours["x_final"], ref["x_final"], rtol=1e-4, atol=2e-4)
            _check_reports(reports)
        fast = model.infer_special(free["semantic_tokens"], None, None, **kw)
        assert torch.equal(fast, model.infer_special(free["semantic_tokens"], None, None, **kw))
    finally:
        model.set_low_latency(False)
    assert torch.equal(model.infer_special(free["semantic_tokens"], None, None, **kw), base)
