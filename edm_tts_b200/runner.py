"""Batch-sharded S2A decode over the GPUs of one box: one process per GPU (torchrun), contiguous shard of utterances per
rank, no collective during the decode, one all_gather of the codes at the end (the reference's
own multi-GPU use of this path, utility_scripts/dump_tokens/dump_tokens.py:152-253, shards batches per rank the same way).

Every utterance's decode is independent of the others, and the in-kernel Philox counters are indexed by the global
utterance index, so the gathered result is bit-identical to a single-GPU decode of the whole batch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous balanced split: rank r decodes rows [B*r/R, B*(r+1)/R)."""
    return batch * rank // world, batch * (rank + 1) // world


def _slice_noise(noise, B, T, lo, hi):
    """cat_gumbel [S-1, B*T, V] / remask_gumbel [S-1, B, T] / forced_* [S, B, T] -> the shard's rows."""
    if noise is None:
        return None
    if noise.dim() == 3 and noise.shape[1] == B * T:      # [S-1, B*T, V]
        return noise.view(noise.shape[0], B, T, noise.shape[-1])[:, lo:hi].reshape(noise.shape[0], (hi - lo) * T, noise.shape[-1])
    return noise[:, lo:hi]


class ShardedDecoder:
    """decode_fn(semantic_tokens, acoustic_prompt_tokens, semantic_prompt_tokens, steps=, temperature=, batch_offset=, **kw)
    -> LongTensor [b, Q, T] for the rows it is given (normally InjectionConformerModel.infer_special)."""

    def __init__(self, decode_fn, group=None, gather_dtype=torch.int16, keep_gather_dtype=False):
        self.decode_fn = decode_fn
        self.group = group
        self.keep_gather_dtype = keep_gather_dtype  # return the gathered int16 codes as they are (the on-disk dtype) instead of int64
        self.gather_dtype = gather_dtype  # codes are < 1024; int16 is also the reference's on-disk format (codes_dataset.py:41-42)

    def __call__(self, semantic_tokens, acoustic_prompt_tokens=None, semantic_prompt_tokens=None, steps=1, temperature=1.0, *,
                 seed=0, cat_gumbel=None, remask_gumbel=None, forced_ids=None, forced_masks=None, forced_coarse=None):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        B, T = semantic_tokens.shape
        lo, hi = shard_bounds(B, rank, world)
        kw = dict(steps=steps, temperature=temperature, seed=seed, batch_offset=lo)
        for name, val in (("cat_gumbel", cat_gumbel), ("remask_gumbel", remask_gumbel), ("forced_ids", forced_ids), ("forced_masks", forced_masks)):
            if val is not None:
                kw[name] = _slice_noise(val, B, T, lo, hi)
        if forced_coarse is not None:
            kw["forced_coarse"] = forced_coarse[lo:hi]
        sl = slice(lo, hi)
        ap = None if acoustic_prompt_tokens is None else acoustic_prompt_tokens[sl]
        sp = None if semantic_prompt_tokens is None else semantic_prompt_tokens[sl]
        local = self.decode_fn(semantic_tokens[sl], ap, sp, **kw) if hi > lo else None
        if world == 1:
            return local.to(self.gather_dtype) if self.keep_gather_dtype else local
        # shards differ by at most one row: pad to the largest, gather, trim
        per = max(shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0] for r in range(world))
        ref = local if local is not None else semantic_tokens
        if B >= world:
            Q = local.shape[1]       # every rank has rows: no extra collective, no host synchronisation on the decode path
        else:                        # some ranks are empty and do not know the number of codebooks
            q_t = torch.tensor([local.shape[1] if local is not None else 0], device=ref.device)
            dist.all_reduce(q_t, op=dist.ReduceOp.MAX, group=self.group)
            Q = int(q_t.item())
        buf = torch.zeros(per, Q, T, device=ref.device, dtype=self.gather_dtype)
        if local is not None:
            buf[: hi - lo] = local.to(self.gather_dtype)
        # gathered as raw bytes: neither NCCL nor gloo has an int16 type
        raw = buf.view(torch.uint8)
        raw_parts = [torch.empty_like(raw) for _ in range(world)]
        dist.all_gather(raw_parts, raw, group=self.group)
        parts = [r.view(self.gather_dtype) for r in raw_parts]
        out = []
        for r, part in enumerate(parts):
            a, b = shard_bounds(B, r, world)
            out.append(part[: b - a])
        out = torch.cat(out, dim=0)
        return out if self.keep_gather_dtype else out.to(torch.int64)
