"""CPU restatement of the reference's S2A decode path, DAC RVQ search, DAC conv encoder / decoder and the HuBERT k-means assignment
(s2a.py, rvq.py, dac_encoder.py, dac_decoder.py, kmeans.py).

TEST INFRASTRUCTURE ONLY. Nothing under edm_tts_b200/ imports this package; it is used by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the checker and the CPU baseline.

Parity status: PINNED. tests/golden/make_golden.py imports the unmodified reference from /root/reference (in the build
container), loads the deterministic weights of oracle/weights.py into it, runs it with injected sampling noise and stores
its outputs under tests/golden/*.pt; tests/test_oracle_golden.py checks this restatement against those files (S2A decode and
eval-mode forward: codes / masks bit-exact, logits 1e-4; RVQ codes bit-exact; DAC encoder latents and decoder waveforms 1e-5).
kmeans.py restates a one-line torch expression (cdist + argmax) and is checked against torch itself.
"""
