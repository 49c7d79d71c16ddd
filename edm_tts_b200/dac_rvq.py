"""Mirror of the DAC ResidualVectorQuantize API (edm_tts/models/dac/vector_quantizer.py:127-252) on the fused CUDA search.

Callers in the reference: DAC.encode_to_codes (dac/modeling_dac.py:163-167, used by utility_scripts/dump_tokens/
dump_tokens.py:213-215 through AudioTokenizer.compute_codes_batch) and DAC.codes_to_features* (modeling_dac.py:173-182).
`install(dac)` swaps the quantizer of a reference DAC module for this one, leaving its conv encoder / decoder untouched.
"""
from __future__ import annotations

import torch

from . import _lib as L
from .weights import pack_rvq_weights


class ResidualVectorQuantize:
    def __init__(self, state_dict: dict, n_codebooks: int = 12, codebook_size: int = 1024, codebook_dim: int = 8,
                 input_dim: int = 1024, prefix: str = "", device="cuda"):
        if not torch.cuda.is_available():
            raise L.EdmError("edm_tts_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        if (codebook_size, codebook_dim, input_dim) != (1024, 8, 1024) or not 1 <= n_codebooks <= 12:
            raise ValueError("the fused RVQ kernel is specialised for 1024 codes x 8 dims, latent 1024, <= 12 codebooks")
        self.n_codebooks, self.codebook_size, self.codebook_dim, self.input_dim = n_codebooks, codebook_size, codebook_dim, input_dim
        self.device = torch.device(device)
        self._t = pack_rvq_weights(state_dict, n_codebooks, prefix, self.device)
        self.training = False

    def eval(self):
        return self

    @torch.no_grad()
    def encode(self, z: torch.Tensor, n_quantizers=None, forced_codes=None, return_latents=False):
        """z [B, 1024, T] (fp32 or bf16) -> codes int64 [B, n_codebooks, T] (csrc/rvq_tc.cuh)."""
        if z.dim() != 3 or z.shape[1] != self.input_dim:
            raise ValueError(f"z must be [B, {self.input_dim}, T]")
        # the tcgen05 path reads z through TMA boxes: rows (the time axis) must be 16-byte multiples
        if z.dtype not in (torch.float32, torch.bfloat16):
            z = z.float()
        z = z.to(self.device)
        B, _, T = z.shape
        al = 8 if z.dtype == torch.bfloat16 else 4
        Tp = (T + al - 1) // al * al
        z = torch.nn.functional.pad(z, (0, Tp - T)) if Tp != T else z.contiguous()
        t = self._t
        codes = torch.empty(B, self.n_codebooks, Tp, device=self.device, dtype=torch.int64)
        lat = torch.empty(B, 96, Tp, device=self.device, dtype=torch.float32) if return_latents else None
        fc = None
        if forced_codes is not None:
            fc = forced_codes.to(self.device, torch.int64)
            fc = torch.nn.functional.pad(fc, (0, Tp - T)) if Tp != T else fc.contiguous()
        e_ws = torch.empty(B * Tp, 96, device=self.device, dtype=torch.float32)
        L.check(L.lib().edm_rvq_encode_tc(L.ptr(z), int(z.dtype == torch.bfloat16), B, Tp, self.n_codebooks, L.ptr(t["w_hi"]), L.ptr(t["w_lo"]), L.ptr(t["b_in"]),
                                          L.ptr(t["cb_packed"]), L.ptr(t["g"]), L.ptr(e_ws), L.ptr(codes), L.ptr(fc), L.ptr(lat),
                                          L.stream_ptr()), "rvq_encode_tc")
        if Tp != T:
            codes = codes[:, :, :T].contiguous()
            lat = lat[:, :, :T].contiguous() if lat is not None else None
        return (codes, lat[:, : self.n_codebooks * self.codebook_dim]) if return_latents else codes

    def forward(self, z: torch.Tensor, n_quantizers=None) -> dict:
        """ResidualVectorQuantize.forward, eval mode (vector_quantizer.py:146-210): all codebooks are searched; the
        quantizer-dropout mask only limits which levels are summed into "z"."""
        codes, latents = self.encode(z, return_latents=True)
        n_q = n_quantizers or self.n_codebooks
        use = min(self.n_codebooks, n_q + 1)
        zq = self.from_codes(codes[:, :use])[0]
        zero = torch.zeros((), device=self.device)
        return {"z": zq, "codes": codes, "latents": latents, "vq/commitment_loss": zero, "vq/codebook_loss": zero}

    __call__ = forward

    def _c2f(self, codes, unreduced):
        codes = codes.to(self.device, torch.int64).contiguous()
        B, Lv, T = codes.shape
        if Lv > self.n_codebooks:
            raise ValueError(f"codes have {Lv} levels, quantizer has {self.n_codebooks}")
        shape = (B, Lv, self.input_dim, T) if unreduced else (B, self.input_dim, T)
        out = torch.empty(shape, device=self.device, dtype=torch.float32)
        L.check(L.lib().edm_codes_to_features(L.ptr(codes), L.ptr(self._t["proj"]), L.ptr(out), B, Lv, T, int(unreduced), L.stream_ptr()),
                "codes_to_features")
        return out

    def from_codes(self, codes):
        """vector_quantizer.py:212-232 -> (z_q [B, D, T], None, codes). The per-level 8-d latents are not materialised."""
        return self._c2f(codes, False), None, codes

    def from_codes_unreduced(self, codes):
        """vector_quantizer.py:234-252 -> [B, L, D, T]."""
        return self._c2f(codes, True)


def install(dac, device="cuda") -> ResidualVectorQuantize:
    """Replace `dac.quantizer` (a reference ResidualVectorQuantize module) by the CUDA path, in place."""
    sd = {k: v.detach() for k, v in dac.quantizer.state_dict().items()}
    q = ResidualVectorQuantize(sd, n_codebooks=dac.n_codebooks, codebook_size=dac.codebook_size, codebook_dim=dac.codebook_dim,
                               input_dim=dac.latent_dim, device=device)
    if "quantizer" in getattr(dac, "_modules", {}):
        del dac._modules["quantizer"]  # nn.Module refuses to overwrite a registered child with a plain object
    object.__setattr__(dac, "quantizer", q)
    return q
