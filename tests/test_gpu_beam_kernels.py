"""Kernel-level parity of the beam kernels — (3a) e2e_beam_candidates, (3b) e2e_beam_combine_prune and
e2e_beam_finalize — against the oracle's own per-step pieces (oracle/beam_oracle.py: blend_ctc, add_lm, _expand,
prune; pinned to the live reference by tests/test_oracle_vs_reference.py), through the C ABI.

Two kinds of cases:
* trace replay: the oracle decodes an utterance and records, per (step, parent), the speller / LM logits, the CTC
  candidates and prefix scores it saw (``trace=``).  The device kernels are then teacher-forced with exactly those
  inputs, one step at a time, and must reproduce the oracle's beam after every step (tokens, parents, order, scores),
  its closed hypotheses and its final N-best.  Fixtures with an <eos> bias exercise the threshold branch
  (src/decode.py:232-241), ``min_len`` gating (:167-170) and the final re-ranking (:180-183);
* hand-made states: fewer live hypotheses than the beam, equal sort keys (stable order: parent order, then top-k
  order, src/decode.py:175-176), the <eos> threshold at its boundary, min_len gating.
Integer outputs exact; scores 2e-5 absolute (device log-softmax vs torch's on the CPU).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SC_TOL = 2e-5


def _ops():
    from e2e_asr_pytorch_b200 import ops
    return ops


def _oracle_step(parents, att_logits, lm_logits, cands, psi, beam, ctc_w, lm_w, step, min_len):
    """One decode step of the oracle over given logits: (new live list, hypotheses closed at this step).
    Children carry their parent's index in ``dec_state``."""
    from oracle import beam_oracle as BO
    pool, closed = [], []
    for b, hyp in enumerate(parents):
        score = F.log_softmax(att_logits[b:b + 1], dim=-1)
        att_logp = score.squeeze(0)                                   # a VIEW: aliased by the LM add when CTC is off
        cur = score
        if ctc_w > 0:
            cur = BO.blend_ctc(score, cands[b], psi[b], hyp.ctc_prob, ctc_w)
        if lm_w > 0:
            cur = BO.add_lm(cur, lm_logits[b:b + 1], lm_w)
        top_vals, top_ids = cur.squeeze(0).topk(beam)
        states = np.zeros((len(cands[b]), 1, 2), np.float32) if ctc_w > 0 else None
        cl, ch = BO._expand(hyp, top_ids, top_vals, b, None, None, states, psi[b] if ctc_w > 0 else None,
                            cands[b] if ctc_w > 0 else None, att_logp)
        if cl is not None and step >= min_len:
            closed.append(cl)
        pool.extend(ch)
    return BO.prune(pool, beam), closed


def _load_state(buf, parents, step):
    """Device beam state of utterance 0 := the oracle's live list."""
    n = len(parents)
    buf.n_live[0] = n
    buf.n_active[0] = n
    for b, hyp in enumerate(parents):
        buf.last_tok[0, b] = hyp.ids[-1] if hyp.ids else 0
        buf.last_tok64[0, b] = hyp.ids[-1] if hyp.ids else 0
        buf.prefix_len[0, b] = step
        buf.score_sum[0, b] = float(sum(hyp.scores)) if hyp.scores else 0.0
        buf.ctc_prob[0, b] = float(hyp.ctc_prob)


def _run_step(ops, buf, att_logits, lm_logits, oracle_cands, psi, vocab, step, ctc_w, lm_w, dev):
    n = att_logits.shape[0]
    beam, n_cand = buf.B, buf.C
    att = torch.zeros((beam, vocab), device=dev)
    att[:n] = att_logits.to(dev)
    lm = None
    if lm_w > 0:
        lm = torch.zeros((beam, vocab), device=dev)
        lm[:n] = lm_logits.to(dev)
    ops.beam_candidates(att, 1, beam, vocab, n_cand, buf.n_active, buf.att_stats, buf.cand)
    stats = buf.att_stats.cpu().numpy()
    want_lp = F.log_softmax(att_logits, dim=-1).numpy()
    got_lp = (att_logits.numpy() - stats[:n, 0:1]) - stats[:n, 1:2]
    assert np.abs(got_lp - want_lp).max() < 3e-6
    if ctc_w > 0:
        got_c = buf.cand.cpu().numpy()[:n, :n_cand]
        assert np.array_equal(got_c, np.asarray(oracle_cands, dtype=np.int32)), (step, got_c, oracle_cands)
        buf.psi[:n, :n_cand] = torch.as_tensor(np.asarray(psi, dtype=np.float32)).to(dev)
    ops.beam_combine_prune(buf, att, lm, vocab, step, ctc_w, lm_w, 1.5, n_run=1)


def _check_beam(buf, live, parent_toks, step, ctc_w):
    """Device beam after the step == the oracle's new live list (order included)."""
    n = int(buf.n_live[0])
    assert n == len(live), (step, n, len(live))
    tok = buf.last_tok[0].cpu().tolist()
    par = buf.parent_slot[0].cpu().tolist()
    ssum = buf.score_sum[0].cpu().numpy()
    cprob = buf.ctc_prob[0].cpu().numpy()
    ptok = buf.parent_tok[0].cpu().tolist()
    hsc = buf.hist_score[step, 0].cpu().numpy()
    assert buf.parent_row[0].cpu().tolist()[:n] == par[:n] and buf.last_tok64[0].cpu().tolist()[:n] == tok[:n]
    for k, child in enumerate(live):
        assert tok[k] == child.ids[-1] and par[k] == child.dec_state, (step, k, tok[:n], par[:n], [c.ids[-1] for c in live], [c.dec_state for c in live])
        assert int(buf.prefix_len[0, k]) == step + 1
        assert abs(float(ssum[k]) - float(sum(child.scores))) < SC_TOL * (step + 1)
        assert abs(float(hsc[k]) - float(child.scores[-1])) < SC_TOL
        if ctc_w > 0:
            assert float(cprob[k]) == float(child.ctc_prob)
        assert ptok[k] == parent_toks[child.dec_state]


@pytest.mark.parametrize("beam,eos_bias,blank_bias,ctc_w,lm_w,n_frames,min_ratio",
                         [(4, 3.0, 6.0, 0.5, 0.3, 120, 0.01), (4, 1.0, 6.0, 0.5, 0.0, 92, 0.01), (4, 4.0, 0.0, 0.0, 0.3, 92, 0.01),
                          (4, 4.0, 0.0, 0.0, 0.0, 64, 0.01), (8, 0.0, 0.0, 0.5, 0.5, 92, 0.01), (2, 0.0, 0.0, 0.5, 0.0, 64, 0.01),
                          (4, 4.0, 6.0, 0.5, 0.3, 120, 0.08), (16, 2.0, 3.0, 0.5, 0.3, 148, 0.01)])
def test_beam_kernels_replay_the_oracle_trace(cuda, beam, eos_bias, blank_bias, ctc_w, lm_w, n_frames, min_ratio):
    ops = _ops()
    from oracle import beam_oracle as BO
    from e2e_asr_pytorch_b200 import synth
    vocab = 31
    asr = synth.build_asr(vocab, synth.TINY_ASR_CFG, seed=0, peak=4.0)
    lm = synth.build_lm(vocab, synth.TINY_LM_CFG, seed=1)
    with torch.no_grad():
        asr.decoder.char_trans.bias[1] += eos_bias
        asr.ctc_layer[0].bias[0] += blank_bias
    feat = synth.utterance(3, n_frames)[None]
    trace = []
    with torch.no_grad():
        nbest = BO.decode_utterance(asr, feat, torch.LongTensor([n_frames]), beam, min_ratio, 0.2, lm=lm if lm_w > 0 else None,
                                    lm_weight=lm_w, ctc_weight=ctc_w, trace=trace)
    n_steps = int(np.ceil(n_frames * 0.2))
    min_len = int(np.ceil(n_frames * min_ratio))
    n_cand = int(1.5 * beam) if ctc_w > 0 else 0
    buf = ops.BeamBuffers(1, beam, n_cand, n_steps, torch.tensor([min_len]), torch.tensor([n_steps]), cuda)
    by_step = [[r for r in trace if r.get("step") == s] for s in range(n_steps)]
    prefixes = [[]]
    closed_total = 0
    for s in range(n_steps):
        recs = by_step[s]
        assert int(buf.n_live[0]) == len(recs) and [r["prefix"] for r in recs] == prefixes, s
        att = torch.cat([r["att_logits"] for r in recs]).float()
        lml = torch.cat([r["lm_logits"] for r in recs]).float() if lm_w > 0 else None
        _run_step(ops, buf, att, lml, [r["cands"] for r in recs] if ctc_w > 0 else None,
                  [r["psi"] for r in recs] if ctc_w > 0 else None, vocab, s, ctc_w, lm_w, cuda)
        n = int(buf.n_live[0])
        tok, par = buf.last_tok[0].cpu().tolist()[:n], buf.parent_slot[0].cpu().tolist()[:n]
        prefixes = [prefixes[p] + [t] for t, p in zip(tok, par)]
        if s + 1 < n_steps:
            assert prefixes == [r["prefix"] for r in by_step[s + 1]], (s, prefixes, [r["prefix"] for r in by_step[s + 1]])
        # per-token scores: the winner's blended score in the parent's top-k
        hsc = buf.hist_score[s, 0].cpu().numpy()
        for k in range(n):
            r = recs[par[k]]
            idx = r["top_ids"].tolist().index(tok[k])
            assert abs(float(hsc[k]) - float(r["top_vals"][idx])) < SC_TOL, (s, k)
        closed_total = int(buf.fin_count[0])
    assert int(buf.status[0]) == 0
    tok, sc, ln, avg, n = (a.cpu().numpy() for a in ops.beam_finalize(buf))
    want = BO.nbest_as_arrays(nbest)
    assert int(n[0]) == len(want)
    for k, (w_tok, w_sc, w_avg) in enumerate(want):
        m = int(ln[0, k])
        assert tok[0, k, :m].tolist() == w_tok.tolist(), (k, tok[0, k, :m].tolist(), w_tok.tolist())
        assert np.abs(sc[0, k, :m] - w_sc).max() < SC_TOL and abs(float(avg[0, k]) - float(w_avg)) < SC_TOL
    n_closed_ref = sum(1 for w in want if len(w[0]) < n_steps or (w[0][-1] == 1 and False))
    print("replay beam %d ctc %.1f lm %.1f: %d steps, closed on device %d, N-best entries shorter than S %d"
          % (beam, ctc_w, lm_w, n_steps, closed_total, n_closed_ref))
    if eos_bias >= 3.0 and min_ratio < 0.05:
        assert closed_total > 0, "the fixture did not exercise <eos> termination"


def _beam_from(tokens, scores, ctc_prob=0.0):
    from oracle import beam_oracle as BO
    return BO.Beam([torch.tensor(int(t)) for t in tokens], [torch.tensor(float(s), dtype=torch.float32) for s in scores],
                   None, None, None, None, np.float32(ctc_prob))


@pytest.mark.parametrize("case,vocab", [("equal_keys", 31), ("few_live", 31), ("eos_closes", 31), ("eos_boundary_not_closed", 31),
                                        ("min_len_blocks", 31), ("lm_alias", 31), ("all_three", 31),
                                        # 64 / 128-wide rows: 2 / 4 cached entries per lane; wider: the one-pass streaming selection
                                        ("all_three", 50), ("equal_keys", 100), ("all_three", 300), ("eos_closes", 300), ("lm_alias", 10000),
                                        ("all_three", 10000), ("equal_keys", 10000)])
def test_beam_combine_prune_hand_made_states(cuda, case, vocab):
    ops = _ops()
    rng = np.random.default_rng(sum(map(ord, case)) + vocab)
    beam, step = 4, 3
    ctc_w, lm_w, min_len = 0.0, 0.0, 0
    # three tokens per parent; fp32-exact scores so that equal sums ARE equal
    parents = [_beam_from([5, 6, 7], [-0.25, -0.25, -0.5]), _beam_from([5, 6, 8], [-0.5, -0.25, -0.25]),
               _beam_from([9, 6, 7], [-1.0, -0.5, -1.0]), _beam_from([4, 4, 4], [-0.75, -0.75, -0.75])]
    att = torch.from_numpy(rng.standard_normal((beam, vocab)).astype(np.float32) * 3)
    lml = torch.from_numpy(rng.standard_normal((beam, vocab)).astype(np.float32) * 2)
    cands = psi = None
    if case == "equal_keys":
        att[1] = att[0]                                   # same logits, same score sums: every child key of parent 1 ties with parent 0's
    elif case == "few_live":
        parents = parents[:2]
    elif case == "eos_closes":
        att[0, 1] = att[0].max() + 4.0                    # <eos> far above everything: log p(eos) ~ 0 > 1.5 * (negative)
        att[2, 1] = att[2].max() + 4.0
    elif case == "eos_boundary_not_closed":
        # <eos> is in the top-k but log p(eos) <= 1.5 * log p(best other): it stays an ordinary token (src/decode.py:235-248)
        parents = parents[:1]                             # one live parent: all of its top-k children survive the prune
        att[0] = torch.from_numpy(rng.standard_normal(vocab).astype(np.float32) * 0.1 - 10.0)
        att[0, 7], att[0, 1] = 5.0, 1.0
    elif case == "min_len_blocks":
        att[1, 1] = att[1].max() + 4.0
        min_len = step + 1                                # closing is not recorded before min_len, the child is still dropped
    elif case == "lm_alias":
        lm_w = 0.5                                        # CTC off + LM on: the <eos> test sees the LM-blended scores (reference aliasing)
        att[3, 1] = att[3].max() + 1.0
        lml[3, 1] = lml[3].max() + 6.0
    elif case == "all_three":
        ctc_w, lm_w = 0.5, 0.3
    n = len(parents)
    n_cand = int(1.5 * beam) if ctc_w > 0 else 0
    if ctc_w > 0:
        lp = F.log_softmax(att, dim=-1)
        cands = [lp[b].topk(n_cand)[1].tolist() for b in range(n)]
        psi = [(-rng.random(n_cand) * 8 - 1).astype(np.float32) for _ in range(n)]
        for b, hyp in enumerate(parents):
            hyp.ctc_prob = np.float32(-0.5 * (b + 1))
    buf = ops.BeamBuffers(1, beam, n_cand, step + 2, torch.tensor([min_len]), torch.tensor([step + 2]), cuda)
    _load_state(buf, parents, step)
    parent_toks = [h.ids[-1] if h.ids else -1 for h in parents]          # before _expand appends <eos> to a closing parent
    live, closed = _oracle_step(parents, att[:n].clone(), lml[:n].clone(), cands, psi, beam, ctc_w, lm_w, step, min_len)
    _run_step(ops, buf, att[:n], lml[:n] if lm_w > 0 else None, cands, psi, vocab, step, ctc_w, lm_w, cuda)
    _check_beam(buf, live, parent_toks, step, ctc_w)
    assert int(buf.fin_count[0]) == len(closed)
    if closed:
        order = sorted(range(len(closed)), key=lambda i: -float(closed[i].mean_score()))      # the list is kept sorted by mean score (stable)
        fs = buf.fin_sum[0].cpu().numpy()
        for slot, i in enumerate(order):
            assert abs(float(fs[slot]) - float(sum(closed[i].scores))) < SC_TOL * (step + 1)
            assert int(buf.fin_step[0, slot]) == step
    if case == "eos_closes":
        assert len(closed) >= 2
    if case == "eos_boundary_not_closed":
        assert len(closed) == 0 and any(c.ids[-1] == 1 for c in live)
    if case == "min_len_blocks":
        assert len(closed) == 0 and all(not (c.dec_state == 1 and c.ids[-1] == 1) for c in live)
    if case == "equal_keys":
        # stable order: of two equal keys the child of parent 0 comes first
        keys = [float(c.mean_score()) for c in live]
        assert any(keys[i] == keys[i + 1] and live[i].dec_state < live[i + 1].dec_state for i in range(len(live) - 1))
