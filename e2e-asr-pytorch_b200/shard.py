"""Utterance sharding across the GPUs of one box + the single N-best all-gather.

The reference parallelises decode only by fanning single utterances out to
joblib worker processes (``bin/test_asr.py:108-109,138-139``).  The recursion
does not shard inside an utterance (sequential in T, beam coupled every step),
so here whole utterances are dealt to ranks (one process per GPU), every rank
decodes its shard with the batched device beam search, and ONE all-gather of a
packed int32 N-best buffer (NCCL over NVLink/NVSwitch; gloo in the CPU tests)
gives every rank the full result in the original utterance order.  Nothing else
crosses GPUs.
"""
import numpy as np
import torch
import torch.distributed as dist


def utterance_cost(n_frames, max_len_ratio=0.2):
    """Relative decode cost of an utterance: steps x encoder frames ~ L^2."""
    n = np.asarray(n_frames, dtype=np.float64)
    return np.ceil(n * max_len_ratio) * (n / 4.0) + 50.0 * n


def plan_shards(lengths, world_size, max_len_ratio=0.2):
    """Longest-processing-time greedy: returns ``world_size`` index arrays (each sorted by
    descending length) whose summed costs are balanced.  Deterministic, so every rank
    computes the same plan from the same length list."""
    lengths = np.asarray(lengths)
    cost = utterance_cost(lengths, max_len_ratio)
    order = np.lexsort((np.arange(len(lengths)), -cost))          # cost desc, index asc
    load = np.zeros(world_size)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))                                    # ties -> lowest rank
        shards[r].append(int(i))
        load[r] += cost[i]
    return [np.array(s, dtype=np.int64) for s in shards]


def estimate_decode_bytes(n_utts, l_max, vocab, beam, n_cand, feat_dim=160, enc_dim=640, att_dim=300,
                          lm_dim=1024, lm_layers=4, encoder_chunk=128):
    """Upper estimate of the HBM ``BeamDecoder.decode_batch`` needs at its peak for ``n_utts`` utterances padded
    to ``l_max`` input frames (every buffer is sized by the LONGEST utterance of the batch).
    Dominant terms: the frame-major posteriors together with the CTC logits they are made from
    (2 x U x Tmax x Vp fp32 — 170 GB for 2620 utterances with a 10k subword vocabulary), the two prefix-state
    buffers, the encoder output and the attention keys; the encoder's chunk temporaries are a constant."""
    u, t = int(n_utts), (int(l_max) // 4 + 3) // 4 * 4
    vp = (int(vocab) + 3) // 4 * 4
    n = u * beam
    feats = u * int(l_max) * feat_dim * 4
    enc = u * t * enc_dim * 4 * 2                                   # padded output + the packed frames it is scattered from
    keys = u * t * att_dim * 4 * 2                                  # proj_k output and its channel-major copy
    post = 2 * u * t * vp * 4 + u * t * 8                           # logits + posteriors (+ blank running sum)
    states = 2 * u * t * beam * max(1, n_cand) * 8
    attn = 3 * n * t * 4
    lstm = n * (lm_layers * (3 * 2 * lm_dim * 2 + 4 * lm_dim * 4) + 4 * lm_dim * 4) + 3 * n * vp * 4
    hist = 3 * (int(l_max) // 5 + 2) * n * 4
    chunk = min(u, encoder_chunk) * int(l_max) * 40 * 128 * 4 * 3 + (3 << 30) * 2      # VGG activations + unfolded operand + GEMM result
    return int(1.1 * (feats + enc + keys + post + states + attn + lstm + hist + chunk))


def make_batches(indices, lengths, max_utts=512, max_padded_frames=None, max_bytes=None, bytes_fn=None):
    """Split ``indices`` (any order) into length-sorted batches: utterances of similar length
    share a batch so that padding and idle decode steps stay small.  ``max_bytes`` with
    ``bytes_fn(n_utts, l_max) -> bytes`` (e.g. a partial of :func:`estimate_decode_bytes`) also closes a batch
    before its estimated footprint exceeds the budget; a single utterance is never refused."""
    lengths = np.asarray(lengths)
    idx = sorted((int(i) for i in indices), key=lambda i: (-int(lengths[i]), i))
    batches, cur = [], []
    for i in idx:
        longest = int(lengths[cur[0]]) if cur else int(lengths[i])
        if cur and (len(cur) >= max_utts or (max_padded_frames and (len(cur) + 1) * longest > max_padded_frames)
                    or (max_bytes and bytes_fn is not None and bytes_fn(len(cur) + 1, longest) > max_bytes)):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


# ---- packed N-best buffer -----------------------------------------------------------------------
# row = [utt_id, n, len_0..len_{B-1}, avgbits_0..avgbits_{B-1}, tok[B*cap], scorebits[B*cap]]  (int32)
def row_width(beam, cap):
    return 2 + 2 * beam + 2 * beam * cap


def pack_nbest(utt_ids, tok, sc, ln, avg, n, cap, rows):
    """CPU int32 [rows, row_width]; unused rows have utt_id = -1."""
    u, beam, have = tok.shape
    buf = torch.zeros((rows, row_width(beam, cap)), dtype=torch.int32)
    buf[:, 0] = -1
    if u == 0:
        return buf
    buf[:u, 0] = torch.as_tensor(np.asarray(utt_ids), dtype=torch.int32)
    buf[:u, 1] = n.to(torch.int32)
    buf[:u, 2:2 + beam] = ln.to(torch.int32)
    buf[:u, 2 + beam:2 + 2 * beam] = avg.to(torch.float32).contiguous().view(torch.int32)
    w = min(cap, have)
    t = torch.zeros((u, beam, cap), dtype=torch.int32)
    s = torch.zeros((u, beam, cap), dtype=torch.float32)
    t[:, :, :w], s[:, :, :w] = tok[:, :, :w], sc[:, :, :w]
    o = 2 + 2 * beam
    buf[:u, o:o + beam * cap] = t.view(u, -1)
    buf[:u, o + beam * cap:] = s.view(u, -1).view(torch.int32)
    return buf


def unpack_nbest(buf, beam, cap, n_total):
    """Inverse of pack_nbest over the concatenation of all ranks' buffers; returns arrays
    indexed by global utterance id."""
    buf = buf.cpu()
    ids = buf[:, 0].long()
    ok = ids >= 0
    rows, ids = buf[ok], ids[ok]
    assert len(ids) == n_total and len(torch.unique(ids)) == n_total, "N-best gather lost or duplicated utterances"
    order = torch.argsort(ids)
    rows = rows[order]
    o = 2 + 2 * beam
    n = rows[:, 1].clone()
    ln = rows[:, 2:2 + beam].clone()
    avg = rows[:, 2 + beam:o].contiguous().view(torch.float32).clone()
    tok = rows[:, o:o + beam * cap].reshape(n_total, beam, cap).clone()
    sc = rows[:, o + beam * cap:].contiguous().view(torch.float32).reshape(n_total, beam, cap).clone()
    return tok, sc, ln, avg, n


def gather_nbest(local_buf, device=None):
    """The one collective of the sharded decode: all-gather of the packed buffers (same shape
    on every rank).  Returns the [world*rows, width] concatenation on the CPU."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_buf
    world = dist.get_world_size()
    send = local_buf.to(device) if device is not None else local_buf
    recv = torch.empty((world * send.shape[0], send.shape[1]), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send.contiguous())
    return recv.cpu()


def decode_sharded(decode_fn, lengths, beam, max_len_ratio, rank=0, world_size=1, device=None,
                   max_utts=512, max_padded_frames=None):
    """Decode utterances 0..len(lengths)-1 across ``world_size`` ranks.

    ``decode_fn(batch_ids) -> (tok, sc, ln, avg, n)`` CPU tensors for the utterances in
    ``batch_ids`` (what ``BeamDecoder.decode_batch(..., return_arrays=True)`` returns).
    Every rank returns the full (tok, sc, ln, avg, n) arrays in utterance order."""
    lengths = np.asarray(lengths)
    n_total = len(lengths)
    shards = plan_shards(lengths, world_size, max_len_ratio)
    cap = int(np.ceil(lengths.max() * max_len_ratio)) + 1 if n_total else 1
    rows = max(len(s) for s in shards)
    mine = shards[rank]
    parts, ids = [], []
    for batch in make_batches(mine, lengths, max_utts, max_padded_frames):
        parts.append(decode_fn(batch))
        ids.extend(batch)
    if parts:
        width = max(p[0].shape[2] for p in parts)
        pad = lambda a: torch.nn.functional.pad(a, (0, width - a.shape[2]))
        tok = torch.cat([pad(p[0]) for p in parts]); sc = torch.cat([pad(p[1]) for p in parts])
        ln = torch.cat([p[2] for p in parts]); avg = torch.cat([p[3] for p in parts]); n = torch.cat([p[4] for p in parts])
    else:
        tok = torch.zeros((0, beam, 1), dtype=torch.int32); sc = torch.zeros((0, beam, 1))
        ln = torch.zeros((0, beam), dtype=torch.int32); avg = torch.zeros((0, beam)); n = torch.zeros((0,), dtype=torch.int32)
    local = pack_nbest(ids, tok, sc, ln, avg, n, cap, rows)
    return unpack_nbest(gather_nbest(local, device), beam, cap, n_total)
