"""The k-means oracle is the reference's own call (`-torch.cdist(...)` then argmax): check the restated semantics on CPU."""
import torch

from oracle.kmeans import kmeans_assign


def test_kmeans_oracle_semantics():
    g = torch.Generator().manual_seed(0)
    centers = torch.randn(37, 16, generator=g)
    embed = torch.randn(2, 11, 16, generator=g)
    embed[1, 3] = centers[5]
    ids, margin = kmeans_assign(embed, centers, return_margins=True)
    brute = ((embed[:, :, None, :] - centers[None, None]) ** 2).sum(-1).argmin(-1)
    assert torch.equal(ids, brute) and ids[1, 3].item() == 5
    assert (margin >= 0).all()
    # ties go to the first centroid, as torch.argmax does
    dup = torch.cat([centers[:1], centers])
    assert kmeans_assign(centers[None, :1], dup)[0, 0].item() == 0
