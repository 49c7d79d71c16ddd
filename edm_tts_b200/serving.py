"""Low-latency serving helper: one S2A decode shape captured in a CUDA graph and replayed per request.

InjectionConformerModel.infer_special issues no host synchronisation and no raw allocation (the re-masking lengths are computed on
the device), so its ~850 launches can be captured as they are; a replay removes the per-launch stream overhead (B=1, 150 frames,
8 steps: 8.4 ms launched, 6.9 ms replayed). The graph reads its tokens from static tensors that each request is copied into. The
sampling noise of the reference (Categorical.sample and the Gumbel of random_topk_mask, modeling_injection_conformer.py:192,
edm_tts/utils/utils.py:49-60) is drawn per request with torch's device RNG into static tensors the captured kernels read (the
injected-noise inputs of the decode), so requests do not share noise; for shapes where those tensors would be large the in-kernel
Philox stream is used instead, keyed by the captured seed plus a device-resident request counter (edm_s2a_set_seed_buffer) that is
bumped before every replay: scalar kernel arguments are baked into a captured graph, a word in device memory is not.
"""
from __future__ import annotations

import torch


class GraphedDecode:
    def __init__(self, model, batch: int, frames: int, prompt_frames: int = 0, steps: int = 8, temperature: float = 1.0, seed: int = 0,
                 fresh_noise=None, max_noise_bytes: int = 256 << 20, low_latency=None):
        """low_latency: capture the decode in the model's single-utterance mode (InjectionConformerModel.set_low_latency); default: on
        for shapes of at most 512 rows, where it applies."""
        self.model, self.steps, self.temperature, self.seed = model, int(steps), float(temperature), int(seed)
        dev = model.device
        B, T, P, V = batch, frames, prompt_frames, model.num_codevectors
        self.low_latency = bool(B * (T + P) <= 512 if low_latency is None else low_latency)
        prev_mode = getattr(model, "low_latency", False)
        model.set_low_latency(self.low_latency)
        self.sem = torch.zeros(B, T, dtype=torch.long, device=dev)
        self.ap = torch.zeros(B, model.num_quantizers, P, dtype=torch.long, device=dev) if P else None
        self.sp = torch.zeros(B, P, dtype=torch.long, device=dev) if P else None
        noise_bytes = (self.steps - 1) * B * T * (V + 1) * 4
        if fresh_noise is None:
            fresh_noise = noise_bytes <= max_noise_bytes
        self.cat = self.rem = None
        if fresh_noise and self.steps > 1:
            self.cat = torch.empty(self.steps - 1, B * T, V, device=dev)
            self.rem = torch.empty(self.steps - 1, B, T, device=dev)
            self._gen = torch.Generator(device=dev)
            self._gen.manual_seed(self.seed)
            self._refresh_noise()
        # Philox mode: the kernels add this device word to the captured seed; __call__ bumps it per request
        self.seed_word = None
        if self.cat is None and self.steps > 1:
            self.seed_word = torch.zeros(1, dtype=torch.int64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                      # warm-up: binds the workspace, loads the kernels
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.codes = self._run()
        model.set_low_latency(prev_mode)   # the mode is baked into the captured launches

    def _run(self):
        from . import _lib as L

        m = self.model
        if self.seed_word is not None:
            L.check(L.lib().edm_s2a_set_seed_buffer(m._ctx, self.seed_word.data_ptr()), "set_seed_buffer")
        try:
            return m.infer_special(self.sem, self.ap, self.sp, steps=self.steps, temperature=self.temperature, seed=self.seed,
                                   cat_gumbel=self.cat, remask_gumbel=self.rem)
        finally:
            if self.seed_word is not None:
                L.check(L.lib().edm_s2a_set_seed_buffer(m._ctx, None), "set_seed_buffer")

    def _refresh_noise(self):
        # Categorical(logits).sample() == argmax(log p + g) with g = -log E, E ~ Exp(1); Gumbel(0, 1) = -log(-log U)
        self.cat.exponential_(generator=self._gen).log_().neg_()
        tiny = torch.finfo(torch.float32).tiny
        self.rem.uniform_(tiny, 1.0, generator=self._gen).clamp_(max=1.0 - torch.finfo(torch.float32).eps).log_().neg_().log_().neg_()

    @torch.no_grad()
    def __call__(self, semantic_tokens, acoustic_prompt_tokens=None, semantic_prompt_tokens=None, clone: bool = True, request_id=None):
        """Same arguments and result as infer_special for the captured shape; `clone=False` returns the graph's static output tensor.
        Philox mode: request n (counted from 0, or `request_id`) decodes as infer_special(..., seed=seed + n)."""
        if tuple(semantic_tokens.shape) != tuple(self.sem.shape):
            raise ValueError(f"captured for semantic tokens of shape {tuple(self.sem.shape)}")
        if (self.ap is None) != (acoustic_prompt_tokens is None) or (self.sp is None) != (semantic_prompt_tokens is None):
            raise ValueError("prompt presence differs from the captured decode")
        self.sem.copy_(semantic_tokens)
        if self.ap is not None:
            if tuple(acoustic_prompt_tokens.shape) != tuple(self.ap.shape) or tuple(semantic_prompt_tokens.shape) != tuple(self.sp.shape):
                raise ValueError("prompt shape differs from the captured decode")
            self.ap.copy_(acoustic_prompt_tokens)
            self.sp.copy_(semantic_prompt_tokens)
        if self.cat is not None:
            self._refresh_noise()
        if self.seed_word is not None:
            if request_id is not None:
                self.seed_word.fill_(int(request_id))
            self.graph.replay()
            out = self.codes.clone() if clone else self.codes
            if request_id is None:
                self.seed_word.add_(1)
            return out
        self.graph.replay()
        return self.codes.clone() if clone else self.codes
