"""The device-side ragged N-best pack (shard.RaggedPacker -> e2e_nbest_pack_ragged) against the CPU packer of the same layout
(shard.pack_nbest_ragged, itself round-trip tested on the CPU in tests/test_host_logic.py): bit-identical buffers, batches in
any order, and the read-back path of gather_nbest at world size 1."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fake_nbest(ids, lengths, beam, ratio, rng, extra=3):
    caps = np.ceil(np.asarray(lengths)[ids] * ratio).astype(np.int64) + 1
    width = int(caps.max()) + extra                       # a batch's row pitch is its own longest utterance's, not the set's
    u = len(ids)
    tok = torch.zeros((u, beam, width), dtype=torch.int32)
    sc = torch.zeros((u, beam, width), dtype=torch.float32)
    ln = torch.zeros((u, beam), dtype=torch.int32)
    n = torch.zeros((u,), dtype=torch.int32)
    for k in range(u):
        n[k] = int(rng.integers(1, beam + 1))
        for b in range(int(n[k])):
            m = int(rng.integers(1, caps[k] + 1))
            ln[k, b] = m
            tok[k, b, :m] = torch.from_numpy(rng.integers(2, 31, m).astype(np.int32))
            sc[k, b, :m] = torch.from_numpy(-rng.random(m).astype(np.float32) * 5)
    avg = torch.from_numpy(-rng.random((u, beam)).astype(np.float32))
    return tok, sc, ln, avg, n


def test_device_pack_equals_cpu_pack_and_round_trips(cuda):
    from e2e_asr_pytorch_b200 import shard
    rng = np.random.default_rng(5)
    lengths = np.array([40, 80, 64, 120, 44, 200, 52, 96, 160, 72, 48, 300])
    beam, ratio, world = 3, 0.2, 2
    shards = shard.plan_shards(lengths, world, ratio)
    size = shard.ragged_size(shards, lengths, beam, ratio)
    bufs = []
    for r, ids in enumerate(shards):
        ids = [int(i) for i in ids]
        batches = [ids[1::2], ids[0::2]]                 # two batches, not in layout order
        packer = shard.RaggedPacker(ids, lengths, beam, ratio, size, cuda)
        packer.reset()
        parts = []
        for b in batches:
            part = _fake_nbest(np.asarray(b), lengths, beam, ratio, rng)
            parts.append((b, part))
            packer.pack(b, *[a.to(cuda) for a in part])
        got = shard.gather_nbest(packer.buf)             # world size 1: read-back only (pinned)
        assert got.is_pinned() and got.shape == (size,)
        # the CPU packer of the same layout, fed the same rows
        width = max(p[1][0].shape[2] for p in parts)
        pad = lambda a: torch.nn.functional.pad(a, (0, width - a.shape[2]))
        order = [i for b, _ in parts for i in b]
        tok = torch.cat([pad(p[0]) for _, p in parts]); sc = torch.cat([pad(p[1]) for _, p in parts])
        ln = torch.cat([p[2] for _, p in parts]); avg = torch.cat([p[3] for _, p in parts]); n = torch.cat([p[4] for _, p in parts])
        want = shard.pack_nbest_ragged(order, tok, sc, ln, avg, n, ids, lengths, beam, ratio, size)
        assert torch.equal(got.cpu(), want), r
        bufs.append(got.clone())
    tok_all, sc_all, ln_all, avg_all, n_all = shard.unpack_nbest_ragged(torch.cat(bufs), shards, lengths, beam, ratio, size)
    assert int((n_all > 0).sum()) == len(lengths) and int(ln_all.max()) <= int(np.ceil(lengths.max() * ratio)) + 1


def test_packer_reports_missing_utterances(cuda):
    from e2e_asr_pytorch_b200 import shard
    lengths = np.array([40, 80, 64])
    shards = shard.plan_shards(lengths, 1, 0.2)
    size = shard.ragged_size(shards, lengths, 2, 0.2)
    packer = shard.RaggedPacker([int(i) for i in shards[0]], lengths, 2, 0.2, size, cuda)
    packer.reset()                                        # nothing packed: every header says "not decoded"
    with pytest.raises(AssertionError):
        shard.unpack_nbest_ragged(shard.gather_nbest(packer.buf), shards, lengths, 2, 0.2, size)
