"""Timings of the BASELINE.json configurations other than the bench.py headline (which is config 2), one JSON line each:
  C1g  S2A decode B=1,  T=150, S=8            (the CPU-runnable case, here on the GPU: small-M / launch-bound regime)
  C2   S2A decode B=64, T=500, S=8            (same as bench.py, for reference)
  C3   S2A decode B=512, T=500, S=8 on one GPU (decoded in chunks of 64; the per-GPU share of config 3 at N=1)
  C5   long-form B=8, T=1500, P=150, S in {1, 8, 32}
  C4   DAC RVQ encode z [32,1024,3000]
Usage: python tools/bench_configs.py > gpurun_out/configs.jsonl
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from edm_tts_b200 import InjectionConformerModel, ResidualVectorQuantize  # noqa: E402
from edm_tts_b200.config import InjectionConformerConfig  # noqa: E402
from edm_tts_b200.synthetic import OracleConfig, make_inputs, make_quantizer_state_dict, make_state_dict  # noqa: E402


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    cfg = OracleConfig()
    sd = make_state_dict(cfg, 0)
    model = InjectionConformerModel(InjectionConformerConfig(), sd, device="cuda")
    out = []
    for name, B, T, P, S, iters in [("C1g", 1, 150, 0, 8, 10), ("C2", 64, 500, 0, 8, 3), ("C3_one_gpu", 512, 500, 0, 8, 1),
                                    ("C5_S1", 8, 1500, 150, 1, 3), ("C5_S8", 8, 1500, 150, 8, 3), ("C5_S32", 8, 1500, 150, 32, 1),
                                    ("C5_B64_S8", 64, 1500, 150, 8, 1)]:
        inp = make_inputs(B, T, P, 1, cfg, seed=1)
        sem = inp["semantic_tokens"].cuda()
        ap = inp["acoustic_prompt_tokens"].cuda() if P else None
        sp = inp["semantic_prompt_tokens"].cuda() if P else None
        ms = timed(lambda: model.infer_special(sem, ap, sp, steps=S, seed=0), iters, warm=1)
        out.append({"config": name, "B": B, "T": T, "P": P, "steps": S, "ms": ms, "frames_per_s": B * T / (ms * 1e-3)})
        print(json.dumps(out[-1]), flush=True)
    rvq = ResidualVectorQuantize(make_quantizer_state_dict(cfg, 0))
    for B, T in [(32, 3000), (4, 3000), (1, 500)]:
        z = torch.randn(B, 1024, T, device="cuda")
        ms = timed(lambda: rvq.encode(z), 10)
        out.append({"config": "C4_rvq", "B": B, "T": T, "ms": ms, "frames_per_s": B * T / (ms * 1e-3)})
        print(json.dumps(out[-1]), flush=True)
    del rvq, model
    torch.cuda.empty_cache()
    # the callers either side of the S2A decode: audio -> codes (dump_tokens) and codes -> audio (inference.py:49)
    from edm_tts_b200.dac import DAC
    from edm_tts_b200.synthetic import make_dac_state_dict
    dac = DAC(make_dac_state_dict(0))
    for B, secs in [(32, 60), (8, 60), (1, 10)]:
        n = 960160 if secs == 60 else 160000
        audio = (torch.randn(B, 1, n, device="cuda") * 0.3).clamp_(-1, 1)
        ms = timed(lambda: dac.encode_to_codes(audio), 3, warm=1)
        T = dac.encoder.lengths(n)[-1]
        out.append({"config": "C4_encode_to_codes", "B": B, "samples": n, "T": T, "ms": ms, "frames_per_s": B * T / (ms * 1e-3)})
        print(json.dumps(out[-1]), flush=True)
        del audio
    for B, T in [(64, 500), (8, 1500), (1, 150)]:
        codes = torch.randint(0, 1024, (B, 12, T), device="cuda")
        ms = timed(lambda: dac.decode_from_codes(codes), 3, warm=1)
        out.append({"config": "decode_from_codes", "B": B, "T": T, "ms": ms, "frames_per_s": B * T / (ms * 1e-3)})
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
