"""Per-utterance, per-hypothesis restatement of the reference beam search
(TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows ``/root/reference/src/decode.py``:

* ``decode_utterance``      <- ``BeamDecoder.forward``        (decode.py:65-183)
* ``_expand``               <- ``Hypothesis.addTopk``         (decode.py:219-263)
* ``Beam.mean_score``       <- ``Hypothesis.avgScore``        (decode.py:214-217)
* constants ``CTC_BEAM_RATIO=1.5``, ``LOG_ZERO=-1e7``        (decode.py:10-11)
* eos threshold 1.5 on pure attention log-probs              (decode.py:220,236-241)

It drives the acoustic model / LM through the same stateful protocol the
reference uses (``decoder.init_state``, ``attention.reset_mem``,
``asr.set_state``, ``asr.attention(query, enc, len)``, ``asr.decoder(x)``,
``lm(tok, lens, hidden)``), one hypothesis at a time on batch-1 tensors, with
the prefix scorer from ``oracle/ctc_prefix_oracle.py``.  It therefore works with
either the reference's own ``ASR``/``RNNLM`` objects or the same-shaped modules
of the product package, and is checked for exact equality (tokens, per-token
scores, N-best order) against the live reference in
tests/test_oracle_vs_reference.py.

Only the configuration the BASELINE configs use is restated: no embedding
fusion plug-in (``emb_decoder=None``).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ctc_prefix_oracle as cpo

CTC_BEAM_RATIO = 1.5      # decode.py:10
LOG_ZERO = -10000000.0    # decode.py:11
EOS_THRESHOLD = 1.5       # decode.py:220
EOS_ID = 1


class Beam:
    """One partial transcript with everything needed to resume it."""
    __slots__ = ("tokens", "scores", "dec_state", "att_map", "lm_state", "ctc_state", "ctc_prob", "_cum")

    def __init__(self, tokens, scores, dec_state, att_map, lm_state, ctc_state, ctc_prob, cum=None):
        self.tokens, self.scores = tokens, scores            # lists of 0-d tensors (ints / fp32)
        self._cum = cum                                      # (len(scores), their left-to-right sum) if already known
        self.dec_state, self.att_map = dec_state, att_map
        if isinstance(lm_state, tuple):
            lm_state = (lm_state[0].cpu(), lm_state[1].cpu())
        elif lm_state is not None:
            lm_state = lm_state.cpu()
        self.lm_state = lm_state
        self.ctc_state, self.ctc_prob = ctc_state, ctc_prob

    @property
    def ids(self):
        return [int(t) for t in self.tokens]

    # reference name, so results can be compared field by field
    outIndex = ids

    def score_sum(self):
        """python sum(): 0 + s0 + s1 + ... one fp32 add at a time, left to right (decode.py:214-217).  A child's sum is
        its parent's sum plus one more add, so the running value is carried along (same adds, same order, same bits)
        instead of being re-added from scratch for every sort key — the long-form cases would otherwise cost O(S^2)."""
        if self._cum is None or self._cum[0] != len(self.scores):
            self._cum = (len(self.scores), sum(self.scores))
        return self._cum[1]

    def mean_score(self):
        return self.score_sum() / len(self.scores)

    avgScore = mean_score


def _expand(parent, top_ids, top_vals, dec_state, att_map, lm_state, ctc_state, ctc_prob, cands, att_logp):
    """Children of ``parent`` for the tokens in ``top_ids`` (best first).  An
    ``<eos>`` whose attention log-prob beats ``EOS_THRESHOLD`` x the best
    non-special token closes the parent instead of spawning a child; any other
    ``<eos>`` is an ordinary token.  Returns (closed_parent_or_None, children)."""
    children, closing = [], None
    for k in range(top_ids.shape[-1]):
        tok = top_ids[k].item()
        if tok == EOS_ID:
            best_other = att_logp[2:].max().item()
            if att_logp[top_ids[k]].item() > EOS_THRESHOLD * best_other:
                closing = top_vals[k].cpu()
                continue
        toks = parent.tokens[:] + [top_ids[k].cpu()]
        scs = parent.scores[:] + [top_vals[k].cpu()]
        st, pr = None, None
        if ctc_state is not None:
            j = cands.index(tok)            # ValueError if the winner was not a CTC candidate (decode.py:252)
            st, pr = ctc_state[j, :, :], ctc_prob[j]
        cum = (len(scs), (parent.score_sum() + scs[-1]) if parent.scores else (0 + scs[-1]))
        children.append(Beam(toks, scs, dec_state, att_map, lm_state, st, pr, cum))
    if closing is not None:
        parent.tokens.append(torch.tensor(EOS_ID))
        parent.scores.append(closing)
        return parent, children
    return None, children


def blend_ctc(score, cands, psi, parent_psi, ctc_weight):
    """decode.py:134-141: CTC prefix deltas spread over the vocabulary (LOG_ZERO elsewhere), blended with the
    attention log-probs [1,V]; token 0 is blocked after blending.  Returns a NEW tensor (``score`` keeps the pure
    attention values the <eos> test looks at)."""
    delta = torch.FloatTensor(psi - parent_psi).to(score.device)
    spread = torch.zeros_like(score).data.fill_(LOG_ZERO)
    for j, c in enumerate(cands):
        spread[0, c] = delta[j]
    score = (1 - ctc_weight) * score + ctc_weight * spread
    score[0, 0] = LOG_ZERO
    return score


def add_lm(score, lm_out, lm_weight):
    """decode.py:144-151: in-place add of the weighted LM log-probs (lm_out: logits [1,V])."""
    score += lm_weight * lm_out.log_softmax(dim=-1)
    return score


def prune(pool, beam_size):
    """decode.py:175-176: stable sort by mean score, best ``beam_size`` kept; ties keep (parent, rank) order."""
    pool.sort(key=lambda b: b.mean_score(), reverse=True)
    return pool[:beam_size]


def decode_utterance(asr, feat, feat_len, beam_size, min_len_ratio, max_len_ratio,
                     lm=None, lm_weight=0.0, ctc_weight=0.0, trace=None, scorer_cls=None, force=None):
    """Joint CTC/attention(+LM) beam search of ONE utterance on CPU.

    feat [1, L, D] float, feat_len [1] long.  Returns the N-best list of Beam
    objects, best first (decode.py:183).  ``trace`` (a list) receives one dict per
    (step, parent) with the tensors the device kernels are checked against.

    ``force`` (token ids): RESCORING for the tie audit — the search keeps one hypothesis and is made
    to follow these tokens, every score computed by exactly the operations above (candidate pre-prune with
    ``beam_size``'s candidate count, CTC blend, LM add, <eos> threshold); returns [that hypothesis].  A token outside
    the CTC candidates raises the reference's ValueError (decode.py:252): the reference could not have produced it.
    """
    assert feat.shape[0] == 1, "Batchsize == 1 is required for beam search"
    assert asr.enable_att
    use_ctc, use_lm = ctc_weight > 0, lm_weight > 0
    if use_ctc:
        assert asr.ctc_weight > 0, "ASR was not trained with CTC decoder"
        n_cand = int(CTC_BEAM_RATIO * beam_size)
    device = feat.device
    scorer_cls = scorer_cls or cpo.PrefixScorerOracle

    dec_state0 = asr.decoder.init_state(1)
    asr.attention.reset_mem()
    n_in = feat_len.cpu().item()
    max_steps = int(np.ceil(n_in * max_len_ratio))
    min_steps = int(np.ceil(n_in * min_len_ratio))
    keep_att = asr.attention.mode == "loc"

    enc, enc_len = asr.encoder(feat, feat_len)
    scorer, state0 = None, None
    if use_ctc:
        post = F.log_softmax(asr.ctc_layer(enc), dim=-1)
        scorer = scorer_cls(post.detach().cpu().numpy())
        state0 = scorer.init_state()

    live = [Beam([], [], dec_state0, None, None, state0, 0)]
    done, stats = [], {"cand_frames": 0, "steps": max_steps, "enc_frames": int(enc.shape[1])}
    for step in range(max_steps if force is None else len(force)):
        pool = []
        for hyp in live:
            tok_prev = hyp.tokens[-1] if hyp.tokens else 0
            tok_prev = torch.LongTensor([tok_prev]).to(device)
            att_prev = hyp.att_map.to(device) if hyp.att_map is not None else None
            lm_prev = hyp.lm_state
            if isinstance(lm_prev, tuple):
                lm_prev = (lm_prev[0].to(device), lm_prev[1].to(device))
            elif lm_prev is not None:
                lm_prev = lm_prev.to(device)
            asr.set_state(hyp.dec_state, att_prev)

            attn, context = asr.attention(asr.decoder.get_query(), enc, enc_len)
            dec_in = torch.cat([asr.pre_embed(tok_prev), context], dim=-1)
            logits, _ = asr.decoder(dec_in)
            score = F.log_softmax(logits, dim=-1)                 # [1, V]
            att_logp = score.squeeze(0)

            cands = psi = new_state = lm_state = None
            rec = {"step": step, "prefix": hyp.ids, "att_logits": logits.detach().clone()} if trace is not None else None
            if use_ctc:
                _, cand_t = score.squeeze(0).topk(n_cand, dim=-1)
                cands = cand_t.cpu().tolist()
                psi, new_state = scorer.cheap_compute(hyp.ids, hyp.ctc_state, cands)
                stats["cand_frames"] += len(cands) * scorer.input_length
                score = blend_ctc(score, cands, psi, hyp.ctc_prob, ctc_weight)
                if rec is not None:
                    rec.update(cands=list(cands), psi=np.array(psi), parent_psi=np.float32(hyp.ctc_prob),
                               r_prev=np.array(hyp.ctc_state))
            if use_lm:
                lm_out, lm_state = lm(tok_prev.unsqueeze(1), torch.ones([1]), hidden=lm_prev)
                lm_out = lm_out.squeeze(0)
                score = add_lm(score, lm_out, lm_weight)
                if rec is not None:
                    rec["lm_logits"] = lm_out.detach().clone()

            if force is None:
                top_vals, top_ids = score.squeeze(0).topk(beam_size)
            else:
                top_ids = torch.LongTensor([int(force[step])]).to(device)
                top_vals = score.squeeze(0)[top_ids]
            att_map = asr.attention.att_layer.prev_att.cpu() if keep_att else None
            if rec is not None:
                rec.update(top_ids=top_ids.clone(), top_vals=top_vals.clone(), att_logp=att_logp.detach().clone())
                trace.append(rec)
            closed, children = _expand(hyp, top_ids, top_vals, asr.decoder.get_state(), att_map, lm_state,
                                       new_state, psi, cands, att_logp)
            if closed is not None and step >= min_steps:
                done.append(closed)
                if beam_size == 1:
                    return done
            pool.extend(children)
            if force is not None and closed is not None:
                return [closed]
        live = prune(pool, beam_size)

    if force is not None:
        return live
    done += live
    done.sort(key=lambda b: b.mean_score(), reverse=True)
    result = done[:beam_size]
    if trace is not None:
        trace.append({"stats": stats})
    return result


def nbest_as_arrays(nbest):
    """[(token ids, per-token fp32 scores, mean score)] for comparisons/fixtures."""
    out = []
    for b in nbest:
        out.append((np.array(b.ids, dtype=np.int32),
                    np.array([float(s) for s in b.scores], dtype=np.float32),
                    np.float32(float(b.mean_score()))))
    return out
