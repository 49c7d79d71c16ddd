"""GPU parity of the text-to-semantic decode (TextToSemanticWLen.infer through the C ABI) against the
oracle restatement and the reference's golden vectors. Same protocol and tolerances as tests/test_gpu_s2a.py: teacher-forced
upstream decisions, logits max |diff| < 0.15 / mean < 0.02, every discrete disagreement a near-tie of the oracle's own scores."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

MAX_TOL, MEAN_TOL = 0.15, 0.02
_MODELS = {}
TEXTS = {"small": "hello world", "base": "The quick brown fox jumps over the lazy dog.",
         "train": "Grafted onto new rootstock, the old tree bore fruit within two seasons; nobody had expected that."}


def _model(cfg_name, seed=0):
    from edm_tts_b200 import TextToSemanticWLen
    from edm_tts_b200.config import TextToSemanticWLenConfig
    from edm_tts_b200.synthetic import T2SConfig, make_t2s_state_dict
    from tests.golden.make_golden_cfg import T2S_CONFIGS

    key = (cfg_name, seed)
    if key not in _MODELS:
        cfg = T2SConfig(**T2S_CONFIGS[cfg_name])
        sd = make_t2s_state_dict(cfg, seed)
        hcfg = TextToSemanticWLenConfig(hidden_size=cfg.hidden, main_encoder_args=dict(depth=cfg.depth, heads=cfg.heads, ff_mult=4, conv_kernel_size=5),
                                        length_predictor_args=dict(depth=cfg.lp_depth, heads=cfg.lp_heads, ff_mult=4, conv_kernel_size=5))
        _MODELS[key] = (cfg, {k: v.cuda() for k, v in sd.items()}, TextToSemanticWLen(hcfg, sd, max_positions=1024))
    return _MODELS[key]


def _report(reports):
    for r in reports:
        head = f"{r['what']:32s} " + (f"max={r['max']:.4f} mean={r['mean']:.5f} " if "max" in r else " " * 24)
        print(head + f"agree={r['agree']:.5f} mismatch={r['n_mismatch']}/{r['n']} not_near_tie={r['n_not_near_tie']} worst_margin={r['worst_margin']:.4f}")
    for r in reports:
        if "max" in r:
            assert r["max"] < MAX_TOL and r["mean"] < MEAN_TOL, r
        assert r["n_not_near_tie"] == 0, r


@pytest.mark.parametrize("cfg_name,iters,length", [("small", 4, None), ("base", 5, None), ("train", 4, 150), ("base", 1, 40), ("train", 16, 500)])
def test_t2s_teacher_forced_parity_vs_oracle(cfg_name, iters, length):
    from edm_tts_b200.synthetic import make_t2s_noise
    from oracle import t2s as ot2s
    from tests.parity_utils import compare_ids, compare_logits, compare_masks, record

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg, sd, model = _model(cfg_name)
    text = TEXTS[cfg_name]
    case = f"t2s {cfg_name} (hidden {cfg.hidden}, heads {cfg.heads}) iters={iters} length={length}"
    with torch.inference_mode():
        tt = ot2s.text_tokens_of(text, cfg, "cuda")
        ref_len, ref_raw = ot2s.predict_length(sd, cfg, tt, return_raw=True)
    got_len, got_raw = model.predict_length(text)
    print(f"length predictor: oracle {ref_raw.item():.5f} -> {int(ref_len)}, ours {got_raw:.5f} -> {got_len}")
    record(case, what="length predictor (log-length)", oracle=ref_raw.item(), ours=got_raw, oracle_length=int(ref_len), ours_length=got_len)
    assert abs(got_raw - ref_raw.item()) < 2e-2
    # the two lengths may differ only when exp(raw) sits next to an integer boundary
    assert got_len == int(ref_len) or abs(ref_raw.exp().item() - round(ref_raw.exp().item())) < 2e-2 * ref_raw.exp().item()
    length = int(ref_len) if length is None else length
    Lseq = len(text.encode("utf-8")) + length + 4
    noise = make_t2s_noise(Lseq, iters, cfg, seed=5 + Lseq)
    tr = {}
    with torch.inference_mode():
        ref_tokens = ot2s.infer(sd, cfg, tt, pred_iters=iters, gt_length=length, cat_gumbel=noise["cat_gumbel"].cuda(),
                                remask_gumbel=noise["remask_gumbel"].cuda(), trace=tr)
    forced = dict(forced_ids=torch.stack(tr["step_ids"]))
    if iters > 1:
        forced["forced_masks"] = torch.stack(tr["step_masks"])
    ours = model.decode_trace(text, pred_iters=iters, gt_length=length, cat_gumbel=noise["cat_gumbel"], remask_gumbel=noise["remask_gumbel"], **forced)
    torch.cuda.synchronize()
    assert torch.equal(ours["input_ids"], tr["input_ids"]) and torch.equal(ours["full_mask"], tr["full_mask"])
    reports = []
    for i in range(iters):
        last = i == iters - 1
        reports.append(compare_logits(ours["step_logits"][i], tr["step_logits"][i], f"logits iteration {i}"))
        scores = tr["step_logits"][i] if last else tr["step_logits"][i] + noise["cat_gumbel"][i].cuda().view(1, Lseq, -1)
        reports.append(compare_ids(ours["step_ids"][i], tr["step_ids"][i], scores, f"own sampled ids iteration {i}"))
        if not last:
            reports.append(compare_masks(ours["step_masks_raw"][i], tr["step_masks"][i], tr["step_conf"][i], tr["step_cut"][i], f"own re-masking iteration {i}"))
    for r in reports:
        record(case, **r)
    _report(reports)
    # with every upstream decision forced, the emitted tokens are the forced last-iteration ids on the speech positions
    assert torch.equal(ours["tokens"], torch.stack(tr["step_ids"])[-1][tr["full_mask"]])
    # and infer() itself (one C call) reproduces the staged run
    out = model.infer(text, pred_iters=iters, gt_length=length, cat_gumbel=noise["cat_gumbel"], remask_gumbel=noise["remask_gumbel"], **forced)
    assert torch.equal(out.speech_pred_tokens, ours["tokens"]) and out["speech_pred_tokens"].dtype == torch.int64
    assert out.speech_pred_tokens.shape == ref_tokens.shape


@pytest.mark.parametrize("name", ["small_s1", "base_s3", "train_s4", "small_s4"])
def test_t2s_against_reference_golden(name, golden_dir):
    """Free-running CUDA decode vs the unmodified reference (fp32 CPU): the predicted length must match; a 1-iteration decode (no
    cascade) must agree except at near-ties of the reference's own logits; in a multi-iteration run of a random-init network one
    flipped sample changes every later iteration (measured agreement with the reference: 3-60 %), so there the graded quantity is
    the re-masking schedule -- the number of tokens still masked after every iteration equals the reference's. The strict
    multi-iteration check is the teacher-forced test above."""
    from edm_tts_b200.synthetic import make_t2s_noise
    from tests.parity_utils import NEAR_TIE_EPS

    g = torch.load(os.path.join(golden_dir, f"t2s_{name}.pt"))
    cfg, sd, model = _model(g["cfg_name"], g["weight_seed"])
    if g["gt_length"] is None:
        length, raw = model.predict_length(g["text"])
        assert abs(raw - g["raw_log_length"]) < 2e-2 and length == g["length"]
    Lseq = len(g["text"].encode("utf-8")) + g["length"] + 4
    noise = make_t2s_noise(Lseq, g["pred_iters"], cfg, seed=g["noise_seed"])
    out = model.infer(g["text"], pred_iters=g["pred_iters"], gt_length=g["gt_length"], cat_gumbel=noise["cat_gumbel"], remask_gumbel=noise["remask_gumbel"])
    tokens, ref = out.speech_pred_tokens.cpu(), g["tokens"].long()
    assert tokens.shape == ref.shape and tokens.min() >= 0 and tokens.max() < 1024
    agree = (tokens == ref).float().mean().item()
    print(name, "agreement with the reference tokens:", round(agree, 3))
    if g["pred_iters"] == 1:
        start = len(g["text"].encode("utf-8")) + 3
        margin = g["last_margin"].float()[start:start + g["length"]]
        assert (margin[tokens != ref] < NEAR_TIE_EPS).all()
        assert agree > 0.9
    else:
        tr = model.decode_trace(g["text"], pred_iters=g["pred_iters"], gt_length=g["gt_length"], cat_gumbel=noise["cat_gumbel"], remask_gumbel=noise["remask_gumbel"])
        assert [int(m.sum()) for m in tr["step_masks"]] == [int(m.sum()) for m in g["masks"]]
        assert torch.equal(tr["tokens"].cpu(), tokens)


def test_t2s_api_and_determinism():
    from oracle import t2s as ot2s

    cfg, sd, model = _model("base")
    text = TEXTS["base"]
    a = model.infer(text, pred_iters=6, seed=3).speech_pred_tokens
    b = model.infer(text, pred_iters=6, seed=3).speech_pred_tokens
    c = model.infer(text, pred_iters=6, seed=4).speech_pred_tokens
    assert torch.equal(a, b) and not torch.equal(a, c) and a.min() >= 0 and a.max() < 1024
    assert a.shape[0] == model.predict_length(text)[0]
    assert model.infer(text, pred_iters=2, gt_length=17).speech_pred_tokens.shape == (17,)
    # embeddings_to_logits on explicit embeddings == the oracle's
    tt = ot2s.text_tokens_of(text, cfg, "cuda")
    ids, _ = ot2s.build_sequence(cfg, tt, 33)
    emb = model.input_embedding(ids)
    with torch.inference_mode():
        want = ot2s.embeddings_to_logits(sd, cfg, torch.nn.functional.embedding(ids, sd["input_embedding.weight"]))
    got = model.embeddings_to_logits(emb)
    assert got.shape == want.shape and (got - want).abs().max().item() < MAX_TOL
    with pytest.raises(ValueError):
        model.infer("x" * 2000, gt_length=10)
    with pytest.raises(ValueError):
        model.infer(text, pred_iters=3, gt_length=20, remask_gumbel=torch.zeros(2, 1, 7))
    with pytest.raises(NotImplementedError):
        model.train()
    with pytest.raises(ValueError):
        from edm_tts_b200 import TextToSemanticWLen
        from edm_tts_b200.config import TextToSemanticWLenConfig
        TextToSemanticWLen(TextToSemanticWLenConfig(hidden_size=320), {})
