"""Token shard files written by edm_tts_b200.token_shards are what the reference's readers expect
(utility_scripts/dump_tokens/dump_tokens.py:217-251 -> edm_tts/datasets/codes_dataset.py:68-83)."""
import os

import pytest
import torch

from edm_tts_b200.token_shards import TokenShardWriter, read_token_shards


def test_writer_layout_rollover_and_reader(tmp_path):
    g = torch.Generator().manual_seed(0)
    w = TokenShardWriter(str(tmp_path / "libri_train"), rank=3, max_files_per_output_file=4)
    truth = {}
    for batch in range(3):                      # batches of 2, 3, 2 utterances -> files roll over after >= 4 utterances
        B = (2, 3, 2)[batch]
        T = 40
        ac = torch.randint(0, 1024, (B, 12, T), generator=g)
        sc = torch.randint(0, 1024, (B, T), generator=g)
        lens = [T - 3 * i for i in range(B)]
        ids = [f"utt_{batch}_{i}" for i in range(B)]
        w.add_batch(ids, ac, sc, lens, transcriptions=[f"text {n}" for n in ids])
        for i, n in enumerate(ids):
            truth[n] = (ac[i, :, : lens[i]], sc[i, : lens[i]])
    files = w.close()
    assert [os.path.basename(f) for f in files] == ["3_0.pt", "3_1.pt"]          # 5 utterances, then the remaining 2
    d0 = torch.load(files[0])
    assert list(d0) == ["utt_0_0", "utt_0_1", "utt_1_0", "utt_1_1", "utt_1_2"]
    item = d0["utt_1_2"]
    assert item["acoustic_codes"].shape == (12, 34) and item["semantic_codes"].shape == (34,) and item["transcription"] == "text utt_1_2"
    assert item["acoustic_codes"].dtype == torch.int16
    # the reference reader's view of the same files
    seen = {}
    for key, ex in read_token_shards(str(tmp_path)):
        assert ex["acoustic_tokens"].shape == (ex["length"], 12) and ex["semantic_tokens"].shape == (ex["length"], 1)
        assert ex["acoustic_tokens"].dtype == torch.int16
        seen[key] = ex
    assert len(seen) == 7
    names = list(truth)
    first = seen["3_0_0"]
    assert torch.equal(first["acoustic_tokens"].long(), truth[names[0]][0].transpose(0, 1))
    assert torch.equal(first["semantic_tokens"][:, 0].long(), truth[names[0]][1])


def test_writer_rejects_mismatched_lengths(tmp_path):
    w = TokenShardWriter(str(tmp_path))
    with pytest.raises(ValueError):
        w.add_batch(["a"], torch.zeros(1, 12, 10, dtype=torch.long), torch.zeros(1, 9, dtype=torch.long))
    with pytest.raises(ValueError):
        w.add_batch(["a", "b"], torch.zeros(1, 12, 10, dtype=torch.long), torch.zeros(1, 10, dtype=torch.long))
    assert w.close() == []
