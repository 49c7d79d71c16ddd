"""Timeline of one single-utterance decode from a -DEDM_KTRACE build (tools/gpu_ktrace.sh): per kernel (block 0) the %globaltimer
stamps entry / PDL wait over / [GEMM: first operands landed, last MMA committed, accumulator in registers] / end."""
import ctypes as C
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import edm_tts_b200._lib as L  # noqa: E402

L.LIB_PATH = os.path.abspath(sys.argv[1])
from edm_tts_b200 import InjectionConformerModel  # noqa: E402
from edm_tts_b200.config import InjectionConformerConfig  # noqa: E402
from edm_tts_b200.synthetic import OracleConfig, make_inputs, make_state_dict  # noqa: E402

B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 150)
LOW_LATENCY = len(sys.argv) > 4 and sys.argv[4] == "low_latency"
cfg = OracleConfig()
model = InjectionConformerModel(InjectionConformerConfig(), make_state_dict(cfg, 0), device="cuda")
sem = make_inputs(B, T, 0, 1, cfg, seed=1)["semantic_tokens"].cuda()
model.set_low_latency(LOW_LATENCY)
lib = L.lib()
lib.edm_ktrace_dump.argtypes = [C.c_void_p, C.c_int]
lib.edm_ktrace_dump.restype = C.c_int
for _ in range(3):
    model.infer_special(sem, None, None, steps=8, seed=0)
lib.edm_ktrace_dump(None, 0)
model.infer_special(sem, None, None, steps=8, seed=0)
buf = (C.c_ulonglong * (8 * 4096))()
n = lib.edm_ktrace_dump(buf, 4096)
rows = [[buf[8 * i + j] for j in range(8)] for i in range(min(n, 4096))]
names = {1: "layernorm", 2: "attention", 3: "conv_stream", 4: "layernorm_splitk"}
name = lambda k: names.get(k, f"gemm<{k - 100}>")
print(f"{n} traced launches; first conformer block (times in us relative to the first entry):")
t0 = rows[1][1]
for r in rows[1:16]:
    rel = [(v - t0) / 1e3 if v else None for v in r[1:7]]
    print(f"  {name(r[0]):12s} " + "  ".join(f"{x:8.2f}" if x is not None else "       -" for x in rel))
agg = defaultdict(lambda: defaultdict(list))
for i in range(1, len(rows)):
    r, pr = rows[i], rows[i - 1]
    k = name(r[0])
    agg[k]["wait_over - prev end"].append((r[2] - pr[6]) / 1e3)
    agg[k]["end - wait_over"].append((r[6] - r[2]) / 1e3)
    agg[k]["entry -> wait_over"].append((r[2] - r[1]) / 1e3)
    if r[0] >= 100 and r[3] and r[4] and r[5]:
        agg[k]["first operands - wait_over"].append((r[3] - r[2]) / 1e3)
        agg[k]["last commit - first operands"].append((r[4] - r[3]) / 1e3)
        agg[k]["acc in regs - last commit"].append((r[5] - r[4]) / 1e3)
        agg[k]["end - acc in regs"].append((r[6] - r[5]) / 1e3)
med = lambda v: sorted(v)[len(v) // 2]
for k, d in agg.items():
    print(f"{k} ({len(d['end - wait_over'])} launches): " + "; ".join(f"{kk} {med(v):.2f}" for kk, v in d.items()))
tot = (rows[-1][6] - rows[0][1]) / 1e3
print(f"first entry -> last end: {tot:.1f} us")
