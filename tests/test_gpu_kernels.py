"""Kernel-level parity on the B200: CUDA path (through the C ABI) vs the CPU oracle.

Tolerances: posteriors 2e-6 abs (fp32 log-softmax); prefix scores / states 1e-4 abs or
2 ulp, whichever is larger (tests/_util.prefix_tolerance); integer outputs exact.
"""
import numpy as np
import pytest
import torch

from tests._util import posteriors, assert_prefix_close, LOGZERO

pytestmark = pytest.mark.gpu


def _ops():
    from e2e_asr_pytorch_b200 import ops, _lib
    return ops, _lib


# ----------------------------------------------------------------------------------------------
# (1) CTC posterior + empty-prefix state
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_utts,t_max,vocab", [(3, 17, 31), (2, 9, 5), (4, 33, 200), (2, 6, 10000), (1, 1, 31)])
def test_ctc_log_softmax_matches_torch(cuda, n_utts, t_max, vocab):
    ops, _ = _ops()
    from oracle import ctc_prefix_oracle as O
    g = torch.Generator().manual_seed(vocab + t_max)
    logits = torch.randn(n_utts, t_max, vocab, generator=g) * 3
    enc_len = torch.tensor([max(1, t_max - 2 * i) for i in range(n_utts)], dtype=torch.int32)
    x = ops.ctc_log_softmax(logits.to(cuda), enc_len.to(cuda), apply_relu=True).cpu()
    vp = x.shape[-1]
    assert x.shape == (t_max, n_utts, vp) and vp % 4 == 0 and vp >= vocab
    ref = torch.log_softmax(torch.relu(logits), dim=-1)                       # decode.py:94-95 on the CPU
    for u in range(n_utts):
        n = int(enc_len[u])
        assert torch.allclose(x[:n, u, :vocab], ref[u, :n], atol=2e-6, rtol=0), (x[:n, u, :vocab] - ref[u, :n]).abs().max()
        assert (x[n:, u] == LOGZERO).all() and (x[:, u, vocab:] == LOGZERO).all()
    r0 = ops.ctc_init_state(x.to(cuda), enc_len.to(cuda)).cpu().numpy()
    for u in range(n_utts):
        n = int(enc_len[u])
        want = O.blank_state(x[:n, u, :vocab].numpy())                         # same device posteriors -> bit exact
        assert np.array_equal(r0[u, :n, 0], want)


def test_ctc_log_softmax_without_relu(cuda):
    ops, _ = _ops()
    logits = torch.randn(2, 5, 31, generator=torch.Generator().manual_seed(3))
    x = ops.ctc_log_softmax(logits.to(cuda), None, apply_relu=False).cpu()
    assert torch.allclose(x[:, :, :31].transpose(0, 1), torch.log_softmax(logits, -1), atol=2e-6, rtol=0)


# ----------------------------------------------------------------------------------------------
# (2) prefix score: a beam-search-shaped chain of steps against the oracle
# ----------------------------------------------------------------------------------------------
def _chain(cuda, rng, n_utts, t_lens, vocab, beam, n_cand, n_steps, flags=0, check_states=True):
    ops, L = _ops()
    ulps = 8.0 if (flags & L.PREFIX_FAST_MATH) else 2.0      # MUFU fast math is opt-in and looser
    from oracle import c_oracle as CO
    t_max = max(t_lens)
    post = posteriors(rng, n_utts, t_max, vocab)                               # [U,T,V]
    vp = ops.padded_vocab(vocab)
    x = np.full((t_max, n_utts, vp), LOGZERO, np.float32)
    x[:, :, :vocab] = post.transpose(1, 0, 2)
    for u, n in enumerate(t_lens):
        x[n:, u] = LOGZERO
    xd = torch.from_numpy(x).to(cuda)
    enc_len = torch.tensor(t_lens, dtype=torch.int32, device=cuda)
    r_prev_d = ops.ctc_init_state(xd, enc_len)
    # oracle-side beams: per utterance a list of (prefix, state[T,2])
    beams = [[([], CO.blank_state(post[u, :t_lens[u]]))] for u in range(n_utts)]
    prev_lane = torch.zeros(n_utts * beam, dtype=torch.int32)
    status = torch.zeros(n_utts, dtype=torch.int32, device=cuda)
    worst = 0.0
    r_bufs = [torch.empty((n_utts, t_max, beam * n_cand, 2), device=cuda) for _ in range(2)]
    for step in range(n_steps):
        n_live = torch.tensor([len(b) for b in beams], dtype=torch.int32)
        cand = np.zeros((n_utts, beam, n_cand), np.int32)
        last = np.zeros((n_utts, beam), np.int32)
        plen = np.zeros((n_utts, beam), np.int32)
        for u in range(n_utts):
            for b, (g, _) in enumerate(beams[u]):
                cs = rng.permutation(vocab)[:n_cand].astype(np.int32)
                if (step + b) % 2 == 0 and 1 not in cs:
                    cs[int(rng.integers(n_cand))] = 1                          # exercise the <eos> override
                if g and (step + b) % 3 == 0 and g[-1] not in cs:
                    cs[int(rng.integers(n_cand))] = g[-1]                      # exercise the repeated-token column
                if len(set(cs.tolist())) < n_cand:
                    cs = rng.permutation(vocab)[:n_cand].astype(np.int32)
                cand[u, b], last[u, b], plen[u, b] = cs, (g[-1] if g else 0), len(g)
        r_out = r_bufs[step % 2]
        r_out.fill_(float("nan"))
        psi, _ = ops.ctc_prefix_score(xd, vocab, enc_len, r_prev_d, prev_lane.to(cuda),
                                      torch.from_numpy(last.reshape(-1)).to(cuda), torch.from_numpy(plen.reshape(-1)).to(cuda),
                                      n_live.to(cuda), torch.from_numpy(cand.reshape(-1, n_cand)).to(cuda),
                                      beam, n_cand, flags, r_out=r_out, status=status)
        psi = psi.cpu().numpy().reshape(n_utts, beam, n_cand)
        r_host = r_out.cpu().numpy().reshape(n_utts, t_max, beam, n_cand, 2) if check_states else None
        new_beams, new_lane = [], np.zeros((n_utts, beam), np.int32)
        for u in range(n_utts):
            n = t_lens[u]
            outs = []
            for b, (g, st) in enumerate(beams[u]):
                p_o, r_o = CO.extend(post[u, :n], len(g), g[-1] if g else 0, st, cand[u, b].tolist())
                worst = max(worst, assert_prefix_close(psi[u, b], p_o, "psi step %d utt %d slot %d" % (step, u, b), ulps))
                if check_states:
                    got = r_host[u, :n, b].transpose(1, 0, 2)                  # [C,T,2]
                    if flags & L.PREFIX_SKIP_DEAD_ROWS:
                        first = max(1, len(g)) if g else 0      # row 0 of an empty-prefix extension is always written
                        got, r_cmp = got[:, first:], r_o[:, first:]
                    else:
                        r_cmp = r_o
                    worst = max(worst, assert_prefix_close(got, r_cmp, "r step %d utt %d slot %d" % (step, u, b), ulps))
                outs.append((g, r_o))
            # survivors: random (parent, candidate) pairs, at most `beam`
            picks = [(b, j) for b in range(len(outs)) for j in range(n_cand)]
            keep = [picks[i] for i in rng.permutation(len(picks))[:beam]]
            nb = []
            for k, (b, j) in enumerate(keep):
                g, r_o = outs[b]
                nb.append((g + [int(cand[u, b, j])], np.ascontiguousarray(r_o[j])))
                new_lane[u, k] = b * n_cand + j
            new_beams.append(nb)
        beams, prev_lane, r_prev_d = new_beams, torch.from_numpy(new_lane.reshape(-1)), r_out
    assert int(status.abs().sum()) == 0
    return worst


def _math_flag(L, mode):
    return {"lut": 0, "mufu": L.PREFIX_FAST_MATH, "libm": L.PREFIX_LIBM_MATH}[mode]


@pytest.mark.parametrize("mode", ["lut", "mufu", "libm"])
def test_prefix_score_cfg1_shape(cuda, mode):
    _, L = _ops()
    rng = np.random.default_rng(11)
    w = _chain(cuda, rng, n_utts=3, t_lens=[250, 249, 100], vocab=31, beam=2, n_cand=3, n_steps=12, flags=_math_flag(L, mode))
    print("cfg1-shape chain: max |gpu-oracle| = %.3g (math=%s)" % (w, mode))


@pytest.mark.parametrize("mode", ["lut", "mufu", "libm"])
def test_prefix_score_cfg2_shape(cuda, mode):
    _, L = _ops()
    rng = np.random.default_rng(12)
    w = _chain(cuda, rng, n_utts=5, t_lens=[180, 37, 96, 64, 181], vocab=31, beam=8, n_cand=12, n_steps=20, flags=_math_flag(L, mode))
    print("cfg2-shape chain: max |gpu-oracle| = %.3g (math=%s)" % (w, mode))


def test_prefix_score_row_copy_staging(cuda):
    """E2E_PREFIX_ROW_COPIES: the per-row bulk-copy staging of the posterior tiles gives the same result as the
    tensor-map box copy (both against the oracle)."""
    _, L = _ops()
    rng = np.random.default_rng(12)
    _chain(cuda, rng, n_utts=5, t_lens=[180, 37, 96, 64, 181], vocab=31, beam=8, n_cand=12, n_steps=8, flags=L.PREFIX_ROW_COPIES)
    _chain(cuda, rng, n_utts=3, t_lens=[50, 41, 7], vocab=200, beam=3, n_cand=4, n_steps=5, flags=L.PREFIX_ROW_COPIES)


def test_prefix_score_skip_dead_rows(cuda):
    _, L = _ops()
    rng = np.random.default_rng(13)
    _chain(cuda, rng, n_utts=2, t_lens=[70, 33], vocab=31, beam=4, n_cand=6, n_steps=30, flags=L.PREFIX_SKIP_DEAD_ROWS)


def test_prefix_score_longform_beam16(cuda):
    rng = np.random.default_rng(14)
    w = _chain(cuda, rng, n_utts=2, t_lens=[875, 640], vocab=31, beam=16, n_cand=24, n_steps=6)
    print("cfg4-shape chain: max |gpu-oracle| = %.3g" % w)


def test_prefix_score_large_vocab_gather(cuda):
    rng = np.random.default_rng(15)
    w = _chain(cuda, rng, n_utts=2, t_lens=[60, 45], vocab=10000, beam=8, n_cand=12, n_steps=6)
    print("cfg3-shape chain (gather variant): max |gpu-oracle| = %.3g" % w)


@pytest.mark.parametrize("vocab,beam,n_cand,flags", [(31, 8, 12, 0), (31, 8, 12, "skip"), (31, 16, 24, 0), (300, 8, 12, 0)])
def test_prefix_score_machine_filling_launch(cuda, vocab, beam, n_cand, flags):
    """>= 2 CTAs per SM: the kernel then runs its wide-CTA layout (up to 128 lanes per CTA; small launches are
    split into one-warp CTAs), incl. the fixed-shape (Vp=32, 96 lanes) specialisation and the gather variant."""
    _, L = _ops()
    rng = np.random.default_rng(17)
    n_utts = 320
    t_lens = [int(t) for t in rng.integers(9, 70, n_utts)]
    t_lens[0], t_lens[1] = 97, 5
    fl = L.PREFIX_SKIP_DEAD_ROWS if flags == "skip" else 0
    w = _chain(cuda, rng, n_utts=n_utts, t_lens=t_lens, vocab=vocab, beam=beam, n_cand=n_cand, n_steps=4, flags=fl)
    print("machine-filling launch V=%d B=%d: max |gpu-oracle| = %.3g" % (vocab, beam, w))


@pytest.mark.parametrize("beam,n_cand,flags", [(8, 12, 0), (8, 12, "skip"), (4, 6, "skip")])
def test_prefix_score_bench_sized_launch(cuda, beam, n_cand, flags):
    """>= 1000 CTAs in one launch: the grid size from which the 16-frame-tile instantiations are taken
    (prefix_score.cu: E2E_PREFIX_SMALL_TILE_FROM) — the fixed-shape <32,96,tensor-map,16> kernel that the bench's
    machine-filling launches run (beam 8) and the generic 16-frame-tile kernel (beam 4)."""
    _, L = _ops()
    rng = np.random.default_rng(23)
    n_utts = 1100 if beam == 8 else 1300
    t_lens = [int(t) for t in rng.integers(5, 48, n_utts)]
    t_lens[0], t_lens[1], t_lens[2] = 83, 5, 16
    fl = L.PREFIX_SKIP_DEAD_ROWS if flags == "skip" else 0
    w = _chain(cuda, rng, n_utts=n_utts, t_lens=t_lens, vocab=31, beam=beam, n_cand=n_cand, n_steps=4, flags=fl)
    print("bench-sized launch (%d CTAs) B=%d: max |gpu-oracle| = %.3g" % (n_utts, beam, w))


def test_prefix_score_midsize_vocab_rows(cuda):
    rng = np.random.default_rng(16)
    _chain(cuda, rng, n_utts=2, t_lens=[50, 41], vocab=200, beam=3, n_cand=4, n_steps=5)


# ----------------------------------------------------------------------------------------------
# (2) fused per-step kernel with lazy state evaluation (e2e_ctc_prefix_step): the same beam-search-shaped chains
# ----------------------------------------------------------------------------------------------
def _chain_lazy(cuda, rng, n_utts, t_lens, vocab, beam, n_cand, n_steps, flags=0, n_run=0, drop_live=False):
    """Per step: psi of every (hypothesis, candidate) against the oracle's cheap_compute, and the state the kernel
    built for every live hypothesis (rows from max(1, len-1) - 1 on) against the oracle's state of that prefix."""
    ops, L = _ops()
    from oracle import c_oracle as CO
    t_max = max(t_lens)
    post = posteriors(rng, n_utts, t_max, vocab)
    vp = ops.padded_vocab(vocab)
    x = np.full((t_max, n_utts, vp), LOGZERO, np.float32)
    x[:, :, :vocab] = post.transpose(1, 0, 2)
    for u, n in enumerate(t_lens):
        x[n:, u] = LOGZERO
    xd = torch.from_numpy(x).to(cuda)
    enc_len = torch.tensor(t_lens, dtype=torch.int32, device=cuda)
    r_init = ops.ctc_init_state(xd, enc_len)                                   # [U,T,1,2]
    beams = [[([], CO.blank_state(post[u, :t_lens[u]]), 0, -1)] for u in range(n_utts)]     # (prefix, state, parent slot, parent tok)
    status = torch.zeros(n_utts, dtype=torch.int32, device=cuda)
    r_bufs = [torch.empty((n_utts, t_max, beam, 2), device=cuda) for _ in range(2)]
    worst = 0.0
    n_chk = n_utts if n_run <= 0 else n_run
    for step in range(n_steps):
        n_live = torch.tensor([len(b) for b in beams], dtype=torch.int32)
        cand = np.zeros((n_utts, beam, n_cand), np.int32)
        last = np.zeros((n_utts, beam), np.int32)
        plen = np.zeros((n_utts, beam), np.int32)
        pslot = np.zeros((n_utts, beam), np.int32)
        ptok = np.full((n_utts, beam), -1, np.int32)
        for u in range(n_utts):
            for b, (g, _, ps, pt) in enumerate(beams[u]):
                cs = rng.permutation(vocab)[:n_cand].astype(np.int32)
                if (step + b) % 2 == 0 and 1 not in cs:
                    cs[int(rng.integers(n_cand))] = 1                          # the <eos> override
                if g and (step + b) % 3 == 0 and g[-1] not in cs:
                    cs[int(rng.integers(n_cand))] = g[-1]                      # a repeated token among the candidates
                if len(set(cs.tolist())) < n_cand:
                    cs = rng.permutation(vocab)[:n_cand].astype(np.int32)
                cand[u, b], last[u, b], plen[u, b], pslot[u, b], ptok[u, b] = cs, (g[-1] if g else 0), len(g), ps, pt
        r_prev = r_init if step <= 1 else r_bufs[(step - 1) % 2]
        r_out = r_bufs[step % 2]
        r_out.fill_(float("nan"))
        to = lambda a: torch.from_numpy(a.reshape(-1)).to(cuda)
        psi, _ = ops.ctc_prefix_step(xd, vocab, enc_len, r_prev, to(pslot), to(last), to(ptok), to(plen), n_live.to(cuda),
                                     torch.from_numpy(cand.reshape(-1, n_cand)).to(cuda), beam, n_cand, flags,
                                     r_out=r_out, status=status, n_run=n_run)
        psi = psi.cpu().numpy().reshape(n_utts, beam, n_cand)
        r_host = r_out.cpu().numpy()                                           # [U,T,B,2]
        new_beams = []
        for u in range(n_utts):
            n = t_lens[u]
            outs = []
            for b, (g, st, _, _) in enumerate(beams[u]):
                p_o, r_o = CO.extend(post[u, :n], len(g), g[-1] if g else 0, st, cand[u, b].tolist())
                if u < n_chk:
                    worst = max(worst, assert_prefix_close(psi[u, b], p_o, "psi step %d utt %d slot %d" % (step, u, b)))
                    if step >= 1:       # the hypothesis' own state, as the kernel rebuilt it from its parent's
                        first = max(1, step - 1) - 1
                        worst = max(worst, assert_prefix_close(r_host[u, first:n, b], st[first:], "state step %d utt %d slot %d" % (step, u, b)))
                outs.append((g, r_o))
            picks = [(b, j) for b in range(len(outs)) for j in range(n_cand)]
            n_keep = beam if not (drop_live and u % 3 == 1) else max(1, beam - 1 - (step % 2))
            keep = [picks[i] for i in rng.permutation(len(picks))[:n_keep]]
            if step % 2 == 1:                                                  # make sure repeated tokens survive sometimes
                for b, (g, _) in enumerate(outs):
                    if g and g[-1] in cand[u, b].tolist():
                        keep[0] = (b, cand[u, b].tolist().index(g[-1]))
                        break
            nb = []
            for b, j in keep:
                g, r_o = outs[b]
                nb.append((g + [int(cand[u, b, j])], np.ascontiguousarray(r_o[j]), b, (g[-1] if g else -1)))
            new_beams.append(nb)
        beams = new_beams
    assert int(status.abs().sum()) == 0
    return worst


@pytest.mark.parametrize("mode", ["lut", "poly", "poly_estrin"])
def test_prefix_step_cfg2_shape(cuda, mode):
    _, L = _ops()
    fl = {"lut": 0, "poly": L.PREFIX_POLY_MATH, "poly_estrin": L.PREFIX_POLY_MATH | L.PREFIX_POLY_ESTRIN}[mode]
    rng = np.random.default_rng(12)
    w = _chain_lazy(cuda, rng, n_utts=5, t_lens=[180, 37, 96, 64, 181], vocab=31, beam=8, n_cand=12, n_steps=24, flags=fl, drop_live=True)
    print("fused step, cfg2 shape: max |gpu-oracle| = %.3g (math=%s)" % (w, mode))


def test_prefix_step_cfg1_shape(cuda):
    rng = np.random.default_rng(11)
    w = _chain_lazy(cuda, rng, n_utts=3, t_lens=[250, 249, 100], vocab=31, beam=2, n_cand=3, n_steps=14)
    print("fused step, cfg1 shape: max |gpu-oracle| = %.3g" % w)


def test_prefix_step_longform_beam16(cuda):
    rng = np.random.default_rng(14)
    w = _chain_lazy(cuda, rng, n_utts=2, t_lens=[875, 640], vocab=31, beam=16, n_cand=24, n_steps=6)
    print("fused step, cfg4 shape: max |gpu-oracle| = %.3g" % w)


def test_prefix_step_large_vocab_gather(cuda):
    rng = np.random.default_rng(15)
    w = _chain_lazy(cuda, rng, n_utts=2, t_lens=[60, 45], vocab=10000, beam=8, n_cand=12, n_steps=6)
    print("fused step, cfg3 shape (gather variant): max |gpu-oracle| = %.3g" % w)


def test_prefix_step_midsize_vocab_and_short_utterances(cuda):
    rng = np.random.default_rng(16)
    _chain_lazy(cuda, rng, n_utts=4, t_lens=[50, 41, 3, 1], vocab=200, beam=3, n_cand=4, n_steps=1)
    _chain_lazy(cuda, rng, n_utts=3, t_lens=[50, 41, 7], vocab=200, beam=3, n_cand=4, n_steps=5)
    _chain_lazy(cuda, rng, n_utts=3, t_lens=[17, 16, 33], vocab=31, beam=4, n_cand=6, n_steps=15, drop_live=True)


@pytest.mark.parametrize("beam,n_cand", [(8, 12), (4, 6)])
def test_prefix_step_bench_sized_launch(cuda, beam, n_cand):
    """>= 296 utterances in one launch: the 16-frame-tile instantiations the bench's machine-filling launches run,
    incl. the fixed-shape (Vp = 32) kernel; n_run < U leaves the tail of the batch untouched."""
    rng = np.random.default_rng(23)
    n_utts = 700
    t_lens = [int(t) for t in rng.integers(5, 48, n_utts)]
    t_lens[0], t_lens[1], t_lens[2] = 83, 5, 16
    w = _chain_lazy(cuda, rng, n_utts=n_utts, t_lens=t_lens, vocab=31, beam=beam, n_cand=n_cand, n_steps=5, n_run=640)
    print("fused step, bench-sized launch B=%d: max |gpu-oracle| = %.3g" % (beam, w))


# ----------------------------------------------------------------------------------------------
# drop-in CTCPrefixScore class (src/ctc.py interface)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("t_len,vocab,n_cand", [(1, 31, 3), (2, 31, 3), (7, 31, 12), (60, 31, 12), (50, 200, 12)])
def test_dropin_scorer_against_oracle(cuda, t_len, vocab, n_cand):
    from e2e_asr_pytorch_b200 import CTCPrefixScore
    from oracle import ctc_prefix_oracle as O
    rng = np.random.default_rng(100 + t_len)
    post = posteriors(rng, 1, t_len, vocab)
    dev = CTCPrefixScore(torch.from_numpy(post).to(cuda))
    ora = O.PrefixScorerOracle(post)
    assert (dev.logzero, dev.blank, dev.eos, dev.odim, dev.input_length) == (ora.logzero, ora.blank, ora.eos, ora.odim, ora.input_length)
    assert np.array_equal(dev.x, ora.x)
    r_dev, r_ora = dev.init_state(), ora.init_state()
    assert isinstance(r_dev, np.ndarray) and r_dev.shape == (t_len, 2) and np.array_equal(r_dev, r_ora)
    g = []
    for step in range(min(t_len + 2, 10)):
        cs = [int(c) for c in rng.permutation(vocab)[:n_cand]]
        if step % 2 == 0 and 1 not in cs:
            cs[0] = 1
        if g and step % 3 == 0 and g[-1] not in cs:
            cs[-1] = g[-1]
        try:
            p_o, s_o = ora.cheap_compute(g, r_ora, cs)
        except IndexError:
            with pytest.raises(IndexError):
                dev.cheap_compute(g, r_dev, cs)
            break
        p_d, s_d = dev.cheap_compute(g, r_dev, cs)
        assert p_d.shape == (n_cand,) and s_d.shape == (n_cand, t_len, 2) and p_d.dtype == np.float32
        assert_prefix_close(p_d, p_o, "cheap psi")
        assert_prefix_close(s_d, s_o, "cheap r")
        pf_o, sf_o = ora.full_compute(g, r_ora)
        pf_d, sf_d = dev.full_compute(g, r_dev)
        assert pf_d.shape == (vocab,) and sf_d.shape == (vocab, t_len, 2)
        assert_prefix_close(pf_d, pf_o, "full psi")
        assert_prefix_close(sf_d, sf_o, "full r")
        k = int(rng.integers(n_cand))
        g = g + [cs[k]]
        r_dev, r_ora = s_d[k], s_o[k]


def test_dropin_scorer_device_state_and_errors(cuda):
    from e2e_asr_pytorch_b200 import CTCPrefixScore, _lib
    post = posteriors(np.random.default_rng(5), 1, 20, 31)
    sc = CTCPrefixScore(torch.from_numpy(post).to(cuda), device_state=True)
    r0 = sc.init_state()
    psi, r = sc.cheap_compute([], r0, [1, 2, 3])
    assert psi.is_cuda and r.is_cuda and tuple(r.shape) == (3, 20, 2)
    psi2, _ = sc.cheap_compute([2], r[1], [1, 2, 5])
    assert torch.isfinite(psi2).all()
    with pytest.raises(_lib.E2EError):
        CTCPrefixScore(torch.from_numpy(post))                                  # CPU tensor: no fallback


def test_prefix_too_long_sets_status(cuda):
    ops, L = _ops()
    post = posteriors(np.random.default_rng(6), 1, 4, 31)
    x = np.full((4, 1, 32), LOGZERO, np.float32)
    x[:, 0, :31] = post[0]
    xd = torch.from_numpy(x).to(cuda)
    enc_len = torch.tensor([4], dtype=torch.int32, device=cuda)
    r0 = ops.ctc_init_state(xd, enc_len)
    status = torch.zeros(1, dtype=torch.int32, device=cuda)
    z = torch.zeros(1, dtype=torch.int32, device=cuda)
    ops.ctc_prefix_score(xd, 31, enc_len, r0, z, z, torch.tensor([6], dtype=torch.int32, device=cuda),
                         torch.ones(1, dtype=torch.int32, device=cuda),
                         torch.tensor([[1, 2, 3]], dtype=torch.int32, device=cuda), 1, 3, 0, status=status)
    assert int(status[0]) & L.STATUS_PREFIX_TOO_LONG


# ----------------------------------------------------------------------------------------------
# f-1: fused location-aware attention step
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_utts,beam,t_len,dim,n_filt,half,e_dim", [(3, 8, 180, 300, 10, 100, 640), (2, 2, 37, 24, 4, 10, 40),
                                                                     (1, 16, 875, 300, 10, 100, 640), (4, 3, 300, 128, 12, 25, 96),
                                                                     (2, 1, 50, 32, 7, 3, 17)])
def test_attention_loc_full_matches_oracle(cuda, n_utts, beam, t_len, dim, n_filt, half, e_dim):
    """Whole attention step (conv + energies + softmax + context): attn within 5e-6 abs, context within 2e-5 abs
    of the fp32 CPU restatement; identical results for every hypotheses-per-CTA grouping; only the first n_run
    utterances are written."""
    ops, _ = _ops()
    from oracle import attention_oracle as AO
    g = torch.Generator().manual_seed(t_len + dim)
    n = n_utts * beam
    key = torch.tanh(torch.randn(n_utts, t_len, dim, generator=g))
    value = torch.randn(n_utts, t_len, e_dim, generator=g)
    query = torch.tanh(torch.randn(n, dim, generator=g))
    enc_len = torch.tensor([max(1, t_len - 7 * i) for i in range(n_utts)], dtype=torch.int32)
    prev = torch.softmax(torch.randn(n, t_len, generator=g) * 2, -1)
    for i in range(n):                                  # previous alignments are zero beyond the utterance
        prev[i, int(enc_len[i // beam]):] = 0
    conv_w = torch.randn(n_filt, 2 * half + 1, generator=g) * 0.3
    w_proj = torch.randn(dim, n_filt, generator=g) / n_filt ** 0.5
    w_e = torch.randn(dim, generator=g) / dim ** 0.5 * 4
    want_a, want_c = AO.loc_attention_full(key, value, query, prev, enc_len, conv_w, w_proj, w_e, 0.25, 0.5, beam)
    dev = lambda t: t.to(cuda)
    outs = []
    for nb in (0, 1, 2, 4):
        got_a, got_c = ops.attention_loc_full(dev(key).transpose(1, 2).contiguous(), dev(value), dev(query), dev(prev), dev(enc_len), dev(conv_w), dev(w_proj),
                                              dev(w_e), 0.25, 0.5, beam, hyps_per_unit=nb)
        outs.append((got_a.cpu(), got_c.cpu()))
    got_a, got_c = outs[0]
    err_a, err_c = (got_a - want_a).abs().max().item(), (got_c - want_c).abs().max().item()
    print("attention full U=%d B=%d T=%d A=%d: max |gpu-oracle| attn %.3g ctx %.3g" % (n_utts, beam, t_len, dim, err_a, err_c))
    assert err_a < 5e-6 and err_c < 2e-5
    for a, c in outs[1:]:
        assert torch.equal(a, got_a) and torch.equal(c, got_c)
    for u in range(n_utts):
        assert (got_a[u * beam:(u + 1) * beam, int(enc_len[u]):] == 0).all()
    if n_utts > 1:                                      # n_run: rows of later utterances stay untouched
        attn = torch.full((n, t_len), -7.0, device=cuda)
        ctx = torch.full((n, e_dim), -7.0, device=cuda)
        ops.attention_loc_full(dev(key).transpose(1, 2).contiguous(), dev(value), dev(query), dev(prev), dev(enc_len), dev(conv_w), dev(w_proj), dev(w_e),
                               0.25, 0.5, beam, n_run=1, attn=attn, ctx=ctx)
        assert torch.equal(attn[:beam].cpu(), got_a[:beam]) and (attn[beam:] == -7.0).all() and (ctx[beam:] == -7.0).all()


# ----------------------------------------------------------------------------------------------
# f-2: LSTM step kernels
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,w,k,off,gather", [(37, 1024, 2048, 1024, True), (5, 940, 1240, 0, False), (3, 300, 1240, 940, True), (2, 7, 9, 1, True)])
def test_lstm_split_rows_is_the_exact_three_piece_split(cuda, n, w, k, off, gather):
    """Pieces bit-equal to the torch restatement bf16(x), bf16(x - a1), bf16(x - a1 - a2); rows gathered through
    the parent index; columns outside the slot untouched."""
    ops, _ = _ops()
    from e2e_asr_pytorch_b200.stepper import _split3
    g = torch.Generator().manual_seed(n + w)
    src = (torch.randn(n + 4, w, generator=g) * 3).to(cuda)
    idx = torch.randint(0, n + 4, (n,), generator=g).to(cuda) if gather else None
    dst = torch.full((n + 1, 3 * k), 7.0, dtype=torch.bfloat16, device=cuda)
    ops.lstm_split_rows(src, idx, n, dst, k, off)
    rows = src.index_select(0, idx) if gather else src[:n]
    want = _split3(rows)
    for p in range(3):
        assert torch.equal(dst[:n, p * k + off:p * k + off + w], want[p])
    mask = torch.ones(3 * k, dtype=torch.bool)
    for p in range(3):
        mask[p * k + off:p * k + off + w] = False
    assert (dst[:n][:, mask.to(cuda)] == 7.0).all() and (dst[n] == 7.0).all()
    back = want[0].float() + (want[1].float() + want[2].float())
    assert (back - rows).abs().max().item() <= 2e-7 * rows.abs().max().item()


@pytest.mark.parametrize("n,d,with_table,with_next", [(33, 1024, True, True), (9, 300, False, False), (4, 12, True, False)])
def test_lstm_cell_matches_torch(cuda, n, d, with_table, with_next):
    """c', h' within 1e-6 abs of a float64 evaluation of the nn.LSTM cell; the split of h' lands in the next
    layer's operand."""
    ops, _ = _ops()
    from e2e_asr_pytorch_b200.stepper import _split3
    g = torch.Generator().manual_seed(d)
    gates = torch.randn(n, 4 * d, generator=g) * 2
    bias = torch.randn(4 * d, generator=g)
    table = torch.randn(5, 4 * d, generator=g) if with_table else None
    tok = torch.randint(0, 5, (n,), generator=g) if with_table else None
    c_prev = torch.randn(n + 3, d, generator=g)
    idx = torch.randint(0, n + 3, (n,), generator=g)
    z = gates.double() + bias.double() + (table.double()[tok] if with_table else 0)
    i, f, gg, o = z.chunk(4, dim=-1)
    c_want = torch.sigmoid(f) * c_prev.double()[idx] + torch.sigmoid(i) * torch.tanh(gg)
    h_want = torch.sigmoid(o) * torch.tanh(c_want)
    dev = lambda t: None if t is None else t.to(cuda)
    c_new, h_new = torch.empty(n, d, device=cuda), torch.empty(n, d, device=cuda)
    k_next = 2 * d
    a_next = torch.zeros(n, 3 * k_next, dtype=torch.bfloat16, device=cuda) if with_next else None
    ops.lstm_cell(dev(gates), dev(bias), dev(c_prev), dev(idx), n, c_new, h_new, table=dev(table), tok=dev(tok),
                  a_next=a_next, k_next=k_next if with_next else 0, off_next=0)
    assert (c_new.cpu().double() - c_want).abs().max().item() < 1e-6
    assert (h_new.cpu().double() - h_want).abs().max().item() < 1e-6
    if with_next:
        want = _split3(h_new)
        for p in range(3):
            assert torch.equal(a_next[:, p * k_next:p * k_next + d], want[p])


def test_fused_lstm_stack_matches_nn_lstm(cuda):
    """Three steps of the device LSTM stack (split kernel + split GEMMs + cell kernel, states read through the
    parents' row index) against torch.nn.LSTM in float64 fed with the same parent permutation."""
    _ops()
    from e2e_asr_pytorch_b200.stepper import _FusedLstm
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    torch.manual_seed(3)
    for in_dim, d, layers, table in [(24, 16, 2, False), (16, 16, 3, True)]:
        rnn = torch.nn.LSTM(in_dim, d, num_layers=layers, batch_first=True)
        emb = torch.randn(7, in_dim)
        ref = torch.nn.LSTM(in_dim, d, num_layers=layers, batch_first=True).double()
        ref.load_state_dict({k: v.double() for k, v in rnn.state_dict().items()})
        n = 10
        with torch.no_grad(), _Fp32Math():
            fused = _FusedLstm(rnn.to(cuda), emb.to(cuda) if table else None)
            fused.start(n, cuda)
            h = torch.zeros(layers, n, d, dtype=torch.float64)
            c = torch.zeros(layers, n, d, dtype=torch.float64)
            g = torch.Generator().manual_seed(5)
            for step in range(3):
                tok = torch.randint(0, 7, (n,), generator=g)
                x = emb[tok] if table else torch.randn(n, in_dim, generator=g)
                top = fused.step(n, x0=None if table else x.to(cuda), tok=tok.to(cuda) if table else None)
                out, (h, c) = ref(x.double()[:, None, :], (h, c))
                assert (top.cpu().double() - out[:, 0]).abs().max().item() < 2e-6
                perm = torch.randint(0, n, (n,), generator=g)           # survivors pick parents
                fused.reorder(perm.to(cuda))
                h, c = h[:, perm], c[:, perm]
                assert (fused.hidden(n).cpu().double() - torch.cat(list(h), dim=1)).abs().max().item() < 2e-6


# ----------------------------------------------------------------------------------------------
# f-4: VGG convolutions as split-bf16 tensor-core GEMMs
# ----------------------------------------------------------------------------------------------
def test_conv_unfold_split_matches_unfold(cuda):
    """The unfolded operand equals torch's unfold of the masked NHWC input, piece by piece (bit exact)."""
    ops, _ = _ops()
    from e2e_asr_pytorch_b200.stepper import _split3
    g = torch.Generator().manual_seed(0)
    n, h, w, c = 3, 9, 5, 8
    x = torch.randn(n, h, w, c, generator=g)
    valid = torch.tensor([9, 4, 6], dtype=torch.int32)
    xm = x.clone()
    for i in range(n):
        xm[i, int(valid[i]):] = 0
    cols = torch.nn.functional.unfold(xm.permute(0, 3, 1, 2), 3, padding=1)          # [N, C*9, H*W], k = c*9 + tap
    want = cols.view(n, c, 9, h * w).permute(0, 3, 2, 1).reshape(n * h * w, 9 * c)      # k = tap*C + c
    k = 9 * c
    for p0, m in [(0, n * h * w), (7, 50)]:
        out = torch.zeros(m, 3 * k, dtype=torch.bfloat16, device=cuda)
        ops.conv3x3_unfold_split(x.to(cuda), valid.to(cuda), p0, m, out)
        pieces = _split3(want[p0:p0 + m].to(cuda))
        for q in range(3):
            assert torch.equal(out[:, q * k:(q + 1) * k], pieces[q])


def test_vgg_split_conv_path_is_fp32_accurate(cuda):
    """VGGFrontEnd.forward_masked_split (NHWC + unfold/split + bf16 GEMMs + bias/ReLU/mask kernel) against a float64
    evaluation of forward_masked: not less accurate than the cuDNN fp32 path, padded rows exactly zero."""
    _ops()
    from e2e_asr_pytorch_b200.model import VGGFrontEnd, reference_init_
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    torch.manual_seed(0)
    vgg = VGGFrontEnd(160)
    vgg.apply(reference_init_)
    for m in vgg.extractor:
        if isinstance(m, torch.nn.Conv2d):
            m.bias.data.normal_(0, 0.1)
    vgg.eval()
    lens = torch.tensor([96, 40, 68])
    feat = torch.randn(3, 96, 160)
    for i, l in enumerate(lens):
        feat[i, int(l):] = 0
    with torch.no_grad():
        want, wl = VGGFrontEnd.forward_masked(vgg.double(), feat.double(), lens)
        vgg.float().to(cuda)
        with _Fp32Math():
            got, gl = vgg.forward_masked_split(feat.to(cuda), lens.to(cuda))
            ref32, _ = vgg.forward_masked(feat.to(cuda), lens.to(cuda))
    assert torch.equal(gl.cpu(), wl)
    err_split = (got.cpu().double() - want).abs().max().item()
    err_cudnn = (ref32.cpu().double() - want).abs().max().item()
    print("vgg split conv: max |split - fp64| = %.3g, max |cudnn fp32 - fp64| = %.3g, scale %.3g" % (err_split, err_cudnn, want.abs().max().item()))
    assert err_split < max(2 * err_cudnn, 2e-6 * want.abs().max().item())
    for i, l in enumerate(lens):
        assert (got[i, int(l) // 4:] == 0).all()


@pytest.mark.parametrize("hidden,in_dim,lens", [(16, 24, [30, 22, 22, 9, 1]), (320, 64, [70, 41] + [33] * 20), (40, 12, [700, 650, 310, 12])])
def test_lstm_sequence_matches_nn_lstm(cuda, hidden, in_dim, lens):
    """Packed bidirectional recurrence (RecurrentLayer.forward_packed: split GEMM + persistent kernel + projection) vs
    the layer's own per-utterance forward in float64; groups of 4 / 8 / 16 utterances are all exercised."""
    _ops()
    from e2e_asr_pytorch_b200.model import RecurrentLayer, Encoder, reference_init_
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    torch.manual_seed(hidden)
    layer = RecurrentLayer(in_dim, "LSTM", hidden, True, 0.0, False, 1, "drop", True)
    layer.apply(reference_init_)
    for n, p in layer.named_parameters():
        if "bias" in n:
            p.data.normal_(0, 0.3)
    layer.eval()
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(l, in_dim, generator=g) for l in lens]
    with torch.no_grad():
        ref = copy_double(layer)
        want = torch.cat([ref(x[None].double(), torch.tensor([x.shape[0]]))[0][0] for x in xs], dim=0)
        layer.to(cuda)
        t_len = torch.tensor(lens)
        off = (torch.cumsum(t_len, 0) - t_len).to(torch.int32).to(cuda)
        first, rows = Encoder._groups(t_len)
        with _Fp32Math():
            got = layer.forward_packed(torch.cat(xs).to(cuda), off, t_len.to(torch.int32).to(cuda),
                                       torch.tensor(first, dtype=torch.int32, device=cuda), torch.tensor(rows, dtype=torch.int32, device=cuda))
    err = (got.cpu().double() - want).abs().max().item()
    print("packed BLSTM H=%d: max |gpu - fp64| = %.3g (groups %s)" % (hidden, err, sorted(set(rows))))
    assert err < 5e-6


def copy_double(m):
    import copy
    return copy.deepcopy(m).double()


def test_packed_encoder_matches_per_utterance_calls(cuda):
    """Encoder.forward_ragged_packed (split-bf16 VGG + packed BLSTM stack) vs one batch-1 float64 call per utterance."""
    _ops()
    from e2e_asr_pytorch_b200 import synth
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    asr = synth.build_asr(31, synth.TINY_ASR_CFG, seed=0)
    enc = asr.encoder
    lens = [96, 92, 64, 40, 40, 12]
    feat, fl = synth.padded_batch(list(range(len(lens))), lens)
    with torch.no_grad():
        ref = copy_double(enc)
        wants = [ref(feat[i:i + 1, :n].double(), fl[i:i + 1])[0][0] for i, n in enumerate(lens)]
        enc.to(cuda)
        enc.split_conv = True
        assert enc.packed_supported()
        with _Fp32Math():
            got, gl = enc.forward_ragged_packed(feat.to(cuda), fl.to(cuda), chunk=4)
    assert gl.cpu().tolist() == [n // 4 for n in lens]
    for i, w in enumerate(wants):
        err = (got[i, :w.shape[0]].cpu().double() - w).abs().max().item()
        assert err < 5e-6, (i, err)
        assert (got[i, w.shape[0]:] == 0).all()


def test_conv1_direct_and_fused_pool_match_torch(cuda):
    """First VGG layer straight from the frames (NHWC out) vs F.conv2d on the masked image; bias+ReLU+mask+2x2
    ceil-mode pooling vs the torch ops, odd sizes included."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(3)
    n, l, cin, f, cout = 3, 11, 4, 7, 8
    feat = torch.randn(n, l + 2, cin * f, generator=g)[:, :l]                    # non-contiguous utterances
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.3
    b = torch.randn(cout, generator=g)
    valid = torch.tensor([11, 5, 8], dtype=torch.int32)
    img = feat.reshape(n, l, cin, f).transpose(1, 2).clone()                     # [N, Cin, L, F]
    for i in range(n):
        img[i, :, int(valid[i]):] = 0
    want = torch.relu(torch.nn.functional.conv2d(img.double(), w.double(), b.double(), padding=1)).permute(0, 2, 3, 1)
    got = ops.conv1_direct(feat.to(cuda), w.to(cuda), b.to(cuda), valid.to(cuda), f).cpu()
    assert (got.double() - want).abs().max().item() < 1e-5
    y = torch.randn(n, l, f, cout, generator=g)
    ref = torch.relu(y + b)
    for i in range(n):
        ref[i, int(valid[i]):] = 0
    ref = torch.nn.functional.max_pool2d(ref.permute(0, 3, 1, 2), 2, stride=2, ceil_mode=True).permute(0, 2, 3, 1)
    out = ops.conv_bias_relu_mask_pool(y.to(cuda), b.to(cuda), valid.to(cuda)).cpu()
    assert out.shape == ref.shape and torch.equal(out, ref.contiguous())
