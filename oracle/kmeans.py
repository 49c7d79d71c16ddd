"""Restatement of the semantic tokenizer's k-means assignment (test infrastructure; see oracle/__init__.py).

Reference: edm_tts/models/audio_tokenizer/semantic_tokenizer_hubert/semantic_tokenizer_hubert.py:74-79 (encode) and :82-90
(encode_batch): the centroids are broadcast over the batch, `dists = -torch.cdist(embed, centers, p=2)`, `dists.argmax(-1)`.
The call is a torch built-in, so the restatement is the call itself; margins are returned for the near-tie protocol.
"""
from __future__ import annotations

import torch


def kmeans_assign(embed: torch.Tensor, cluster_centers: torch.Tensor, return_margins: bool = False):
    """embed [B, N, D], cluster_centers [C, D] -> ids [B, N] (first maximum of -distance)."""
    batched = cluster_centers.unsqueeze(0).expand(embed.shape[0], -1, -1)
    dists = -torch.cdist(embed, batched, p=2)
    ids = dists.argmax(dim=-1)
    if not return_margins:
        return ids
    top2 = dists.topk(2, dim=-1)[0]
    return ids, top2[..., 0] - top2[..., 1]
