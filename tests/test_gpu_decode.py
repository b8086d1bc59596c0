"""End-to-end parity on the B200: batched device beam search (BeamDecoder.decode_batch) vs the
per-hypothesis CPU oracle (oracle/beam_oracle.py, itself pinned to the live reference) on
identical random-init weights and seeded synthetic features.

north_star: 1-best token sequences identical except for documented score ties.  The acoustic
model / LM run in fp32 on both sides but through different libraries (cuBLAS/cuDNN vs CPU
kernels), so per-token scores agree to ~1e-5; a differing sequence is accepted only if the
oracle's own scores of the two sequences are closer than TIE_TOL (a tie at fp32 resolution).
"""
import os
import tempfile

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

SCORE_TOL = 2e-4    # |mean score| agreement for identical sequences
TIE_TOL = 5e-4      # score gap below which two hypotheses count as tied


_CPU = {}


def _cpu_copy(m):
    """The decoder's .to(cuda) moves the shared modules; the oracle needs CPU twins."""
    import copy
    if id(m) not in _CPU:
        _CPU[id(m)] = (m, copy.deepcopy(m).cpu())
    return _CPU[id(m)][1]


def _models(vocab=31, peak=4.0):
    from e2e_asr_pytorch_b200 import synth
    asr = synth.build_asr(vocab, synth.TINY_ASR_CFG, seed=0, peak=peak)
    lm = synth.build_lm(vocab, synth.TINY_LM_CFG, seed=1)
    tmp = tempfile.mkdtemp()
    torch.save({"model": lm.state_dict()}, os.path.join(tmp, "lm.pth"))
    yaml.safe_dump({"model": synth.TINY_LM_CFG}, open(os.path.join(tmp, "lm.yaml"), "w"))
    return asr, lm, os.path.join(tmp, "lm.pth"), os.path.join(tmp, "lm.yaml")


def _oracle_nbest(asr, lm, feat, n, beam, lm_w, ctc_w, max_ratio=0.2):
    from oracle import beam_oracle as BO
    asr, lm = _cpu_copy(asr), _cpu_copy(lm)
    with torch.no_grad():
        nb = BO.decode_utterance(asr, feat[None, :n], torch.LongTensor([n]), beam, 0.01, max_ratio,
                                 lm=lm if lm_w > 0 else None, lm_weight=lm_w, ctc_weight=ctc_w)
    return BO.nbest_as_arrays(nb)


def _compare(dev_nbest, ora_nbest, what, rescore=None):
    """Returns (#identical 1-best, #ties) and asserts everything else.  A different 1-best is only accepted as a TIE BY THE
    ORACLE'S OWN SCORES: the device's sequence is in the oracle's N-best within TIE_TOL of its best, or — when the search
    paths diverged early and it is not in that list — ``rescore(ids)`` (the oracle made to follow the device's sequence,
    beam_oracle.decode_utterance(force=...)) gives it a mean score within TIE_TOL of the oracle's best."""
    same = ties = 0
    d0, o0 = dev_nbest[0], ora_nbest[0]
    if d0.outIndex == o0[0].tolist():
        same = 1
        assert abs(float(d0.avgScore()) - float(o0[2])) < SCORE_TOL, what
        assert np.allclose(np.array([float(s) for s in d0.output_scores]), o0[1], atol=2e-3), what
    else:
        cands = [o for o in ora_nbest if o[0].tolist() == d0.outIndex]
        if cands:
            gap = abs(float(cands[0][2]) - float(o0[2]))
        else:
            assert rescore is not None, what + ": device 1-best %s not in the oracle N-best" % d0.outIndex[:10]
            gap = abs(float(rescore(d0.outIndex)) - float(o0[2]))
            print("%s: device 1-best is not in the oracle N-best; rescored by the oracle it is %.3g from the oracle's best" % (what, gap))
        assert gap < TIE_TOL, what + ": not a tie (oracle score gap %.3g)" % gap
        ties = 1
    return same, ties


@pytest.mark.parametrize("beam,lm_w", [(2, 0.0), (8, 0.5), (4, 0.3)])
def test_decode_batch_matches_oracle(cuda, beam, lm_w):
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, lm, lm_path, lm_cfg = _models()
    lens = [64, 120, 92, 200, 76, 148]
    ids = list(range(len(lens)))
    feat, fl = synth.padded_batch(ids, lens)
    dec = BeamDecoder(asr, None, beam, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=lm_w, ctc_weight=0.5).to(cuda)
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    assert len(out) == len(lens)
    same = ties = 0
    for k, n in enumerate(lens):
        ora = _oracle_nbest(asr, lm, feat[k], n, beam, lm_w, 0.5)
        assert len(out[k]) == len(ora)
        s, t = _compare(out[k], ora, "beam %d lm %.1f utt %d" % (beam, lm_w, k))
        same, ties = same + s, ties + t
    print("beam %d lm_w %.1f: identical 1-best %d/%d, ties %d" % (beam, lm_w, same, len(lens), ties))
    assert same >= len(lens) - 1
    # strict drop-in call: one utterance, batch dimension 1 (decode.py:65-67)
    one = dec(feat[1:2, :lens[1]].to(cuda), fl[1:2].to(cuda))
    assert [h.outIndex for h in one] == [h.outIndex for h in out[1]]
    with pytest.raises(AssertionError):
        dec(feat[:2].to(cuda), fl[:2].to(cuda))


def test_decode_without_ctc_and_without_lm(cuda):
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, lm, lm_path, lm_cfg = _models()
    lens = [64, 100]
    feat, fl = synth.padded_batch([0, 1], lens)
    for lm_w in (0.0, 0.5):
        dec = BeamDecoder(asr, None, 4, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=lm_w, ctc_weight=0.0).to(cuda)
        out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
        for k, n in enumerate(lens):
            ora = _oracle_nbest(asr, lm, feat[k], n, 4, lm_w, 0.0)
            _compare(out[k], ora, "no-ctc lm %.1f utt %d" % (lm_w, k))


@pytest.mark.parametrize("eos_bias,blank_bias,ctc_w,lm_w", [(3.0, 6.0, 0.5, 0.3), (1.0, 6.0, 0.5, 0.0), (4.0, 0.0, 0.0, 0.3), (4.0, 0.0, 0.0, 0.0)])
def test_decode_eos_termination(cuda, eos_bias, blank_bias, ctc_w, lm_w):
    """Random-init weights alone never close a hypothesis: bias the speller
    towards <eos> and the CTC head towards blank so that the threshold branch
    (decode.py:232-241), the closed-hypothesis list and the final re-ranking run; with CTC off
    and the LM on, the reference's aliased eos test is exercised too."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, lm, lm_path, lm_cfg = _models()
    with torch.no_grad():
        asr.decoder.char_trans.bias[1] += eos_bias
        asr.ctc_layer[0].bias[0] += blank_bias
    lens = [64, 120, 92]
    feat, fl = synth.padded_batch([0, 1, 2], lens)
    dec = BeamDecoder(asr, None, 4, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=lm_w, ctc_weight=ctc_w).to(cuda)
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    closed = 0
    for k, n in enumerate(lens):
        ora = _oracle_nbest(asr, lm, feat[k], n, 4, lm_w, ctc_w)
        assert [len(h.outIndex) for h in out[k]] == [len(o[0]) for o in ora], (k, [h.outIndex for h in out[k]], [o[0].tolist() for o in ora])
        _compare(out[k], ora, "eos utt %d" % k)
        closed += sum(1 for o in ora if o[0][-1] == 1 and len(o[0]) < int(np.ceil(n * 0.2)))
    assert closed > 0, "the fixture did not exercise <eos> termination"


def test_decode_crash_envelope_is_reported(cuda):
    """ceil(L*max_len_ratio) > T_enc: every candidate's psi becomes log-zero and the reference
    dies with ValueError (decode.py:252) or IndexError (ctc.py:85) — SURVEY.md §7.2-4."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, _, _, _ = _models()
    feat, fl = synth.padded_batch([0], [64])
    dec = BeamDecoder(asr, None, 2, 0.01, 0.5, ctc_weight=0.5).to(cuda)     # 32 steps > T_enc = 16
    with pytest.raises((ValueError, IndexError)):
        dec.decode_batch(feat.to(cuda), fl.to(cuda))


def test_golden_nbest_from_reference(cuda):
    """Fixtures generated by tools/make_golden.py from the UNMODIFIED reference BeamDecoder."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    path = os.path.join(os.path.dirname(__file__), "golden", "beam_nbest_tiny.npz")
    gold = np.load(path, allow_pickle=False)
    asr, lm, lm_path, lm_cfg = _models()
    same = total = 0
    for case in range(int(gold["n_cases"])):
        beam, lm_w, uid, n = (gold["case%d_%s" % (case, k)].item() for k in ("beam", "lm_w", "utt", "len"))
        feat = synth.utterance(int(uid), int(n))
        dec = BeamDecoder(asr, None, int(beam), 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg,
                          lm_weight=float(lm_w), ctc_weight=0.5).to(cuda)
        out = dec(feat[None].to(cuda), torch.LongTensor([int(n)]).to(cuda))
        ora = []
        for k in range(int(gold["case%d_nbest" % case])):
            ora.append((gold["case%d_tok%d" % (case, k)], gold["case%d_sc%d" % (case, k)], gold["case%d_avg%d" % (case, k)]))
        s, _ = _compare(out, ora, "golden case %d" % case)
        same, total = same + s, total + 1
    assert same >= total - 1


def test_split_gemm_is_fp32_accurate(cuda):
    """The recurrent GEMMs run as a 3-way bf16 split on the tensor cores (stepper.SplitLinear).
    Its error against an fp64 product must not exceed that of cuBLAS' own fp32 GEMM."""
    from e2e_asr_pytorch_b200.stepper import SplitLinear
    g = torch.Generator().manual_seed(0)
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    with _Fp32Math():
        for n, k, m in [(2048, 2048, 4096), (1024, 1240, 1200), (1536, 1024, 4096), (8, 300, 31)]:
            x = (torch.randn(n, k, generator=g) * 3).to(cuda)
            w = (torch.randn(m, k, generator=g) / k ** 0.5).to(cuda)
            b = torch.randn(m, generator=g).to(cuda)
            want = (x.double() @ w.double().t() + b.double())
            e_split = (SplitLinear(w, b)(x).double() - want).abs().max().item()
            e_fp32 = (torch.nn.functional.linear(x, w, b).double() - want).abs().max().item()
            print("split-gemm %dx%dx%d: max err %.3g (cuBLAS fp32 %.3g)" % (n, k, m, e_split, e_fp32))
            assert e_split <= 2.0 * e_fp32 + 1e-6


def test_decode_batch_is_order_independent(cuda):
    """decode_batch sorts by length internally; results come back in the caller's order and do not
    depend on what else is in the batch."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, lm, lm_path, lm_cfg = _models()
    lens = [64, 200, 92, 148, 76]
    feat, fl = synth.padded_batch(list(range(5)), lens)
    dec = BeamDecoder(asr, None, 4, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=0.5, ctc_weight=0.5).to(cuda)
    both = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    for k, n in enumerate(lens):
        alone = dec(feat[k:k + 1, :n].to(cuda), fl[k:k + 1].to(cuda))
        assert [h.outIndex for h in alone] == [h.outIndex for h in both[k]], k


def test_decode_batch_of_no_utterances(cuda):
    """An empty shard (more ranks than utterances) decodes to nothing, in every return form, without a launch."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth, _lib
    asr = synth.build_asr(31, synth.TINY_ASR_CFG, seed=0, peak=4.0)
    dec = BeamDecoder(asr, None, 4, 0.01, 0.2, ctc_weight=0.5).to(cuda)
    feat, fl = torch.zeros(0, 0, synth.FEAT_DIM), torch.zeros(0, dtype=torch.long)
    before = _lib.load().e2e_launch_count()
    assert dec.decode_batch(feat.to(cuda), fl.to(cuda)) == []
    for form in (True, "device"):
        tok, sc, ln, avg, n = dec.decode_batch(feat.to(cuda), fl.to(cuda), return_arrays=form)
        assert tok.shape[:2] == (0, 4) and sc.shape[:2] == (0, 4) and ln.shape == (0, 4) and avg.shape == (0, 4) and n.shape == (0,)
        assert tok.is_cuda == (form == "device")
    assert dec.decode_batch_from_host(feat, fl, cuda) == []
    assert _lib.load().e2e_launch_count() == before and dec.last_stats["utterances"] == 0


def test_decode_dataset_feeds_the_reference_result_files(cuda, tmp_path):
    """f-3: decode_dataset + write_results reproduce the reference's per-utterance flow (bin/test_asr.py:138-156):
    same tuples as one-utterance-per-call decoding, files in the reference's layout, scorable by results.score_file."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth, results as R
    from tests.test_results import CharTok
    asr = synth.build_asr(31, synth.TINY_ASR_CFG, seed=0, peak=4.0)
    dec = BeamDecoder(asr, None, 4, 0.01, 0.2, ctc_weight=0.5).to(cuda)
    lens = [64, 92, 40, 76, 92]
    rng = np.random.default_rng(0)
    samples = [("utt%d" % i, synth.utterance(i, n), [int(t) for t in rng.integers(3, 31, 8)] + [1]) for i, n in enumerate(lens)]
    res = R.decode_dataset(dec, samples, cuda, max_utts=3)              # forces two batches
    assert [r[0] for r in res] == [s[0] for s in samples]
    for (name, hyps, truth), (_, feat, tr) in zip(res, samples):
        one = dec(feat[None].to(cuda), torch.LongTensor([feat.shape[0]]).to(cuda))      # the reference's call shape
        assert hyps == [h.outIndex for h in one] and truth == tr
    best, beam = str(tmp_path / "dev_output.csv"), str(tmp_path / "dev_beam.csv")
    R.init_result_files(best, beam)
    R.write_results(res, CharTok(), best, beam)
    assert R.score_file(best)["utterances"] == len(lens)
    assert R.score_file(beam, beam=True)["rows"] == sum(len(r[1]) for r in res)


@pytest.mark.parametrize("vocab,beam,lm_w,lens,max_ratio", [(300, 8, 0.5, [64, 120, 92], 0.07),      # cfg3-like: Vp > 256 -> column-gather prefix kernel
                                                             (31, 16, 0.3, [240, 200, 96], 0.2)])     # cfg4-like: beam 16 (C = 24, 384 lanes per utterance)
def test_decode_large_vocab_and_wide_beam_match_oracle(cuda, vocab, beam, lm_w, lens, max_ratio):
    """End-to-end decode vs the per-hypothesis CPU oracle on the two shapes the bench does not run: a vocabulary
    wide enough for the gather variant of kernel (2) (with the subword config's max_len_ratio) and beam 16."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, lm, lm_path, lm_cfg = _models(vocab=vocab)
    feat, fl = synth.padded_batch(list(range(len(lens))), lens)
    dec = BeamDecoder(asr, None, beam, 0.01, max_ratio, lm_path=lm_path, lm_config=lm_cfg, lm_weight=lm_w, ctc_weight=0.5).to(cuda)
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    same = ties = 0
    for k, n in enumerate(lens):
        ora = _oracle_nbest(asr, lm, feat[k], n, beam, lm_w, 0.5, max_ratio=max_ratio)
        assert len(out[k]) == len(ora)
        s, t = _compare(out[k], ora, "V %d beam %d utt %d" % (vocab, beam, k))
        same, ties = same + s, ties + t
    print("V %d beam %d: identical 1-best %d/%d, ties %d" % (vocab, beam, same, len(lens), ties))
    assert same >= len(lens) - 1


@pytest.mark.parametrize("vocab,beam,lm_w,lens,lazy", [(31, 4, 0.3, [64, 120, 92, 148], False),      # the eager kernel by choice
                                                        (300, 24, 0.0, [64, 100, 80], True)])         # ... and by routing: beam 24 does not fit one CTA
def test_decode_through_the_eager_prefix_kernel_matches_oracle(cuda, vocab, beam, lm_w, lens, lazy):
    """The decode loop's other prefix path: e2e_ctc_prefix_score (a state for every candidate, round 1's kernel) — what
    ``lazy_prefix = False`` selects and what beams whose lanes do not fit the fused kernel's CTA (B > 20) fall back to."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth, ops
    asr, lm, lm_path, lm_cfg = _models(vocab=vocab)
    feat, fl = synth.padded_batch(list(range(len(lens))), lens)
    dec = BeamDecoder(asr, None, beam, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=lm_w, ctc_weight=0.5).to(cuda)
    dec.lazy_prefix = lazy
    assert lazy == (not ops.prefix_step_supported(vocab, beam, dec.ctc_beam_size))      # either switched off or not supported
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    same = ties = 0
    for k, n in enumerate(lens):
        ora = _oracle_nbest(asr, lm, feat[k], n, beam, lm_w, 0.5)
        assert len(out[k]) == len(ora)
        s, t = _compare(out[k], ora, "eager path V %d beam %d utt %d" % (vocab, beam, k))
        same, ties = same + s, ties + t
    print("eager prefix path, V %d beam %d: identical 1-best %d/%d, ties %d" % (vocab, beam, same, len(lens), ties))
    assert same >= len(lens) - 1


@pytest.mark.parametrize("eos_bias,ctc_w,lm_w,ids", [(0.0, 0.0, 0.0, [0, 1, 2, 3]),      # greedy, closes as soon as <eos> wins (decode.py:169-170)
                                                      (4.0, 0.0, 0.3, [0, 1, 2, 3]),      # <eos> wins before min_len: closed hypothesis dropped, no child -> empty N-best
                                                      (0.0, 0.5, 0.3, [0, 1, 3])])        # one CTC candidate (utterance 2 is in the crash envelope, decode.py:252)
def test_decode_beam_1(cuda, eos_bias, ctc_w, lm_w, ids):
    """beam_size = 1 (one CTC candidate): the reference returns at the first closed hypothesis (decode.py:169-170); the device
    path reaches the same N-best because a closed beam-1 hypothesis leaves no child behind."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, lm, lm_path, lm_cfg = _models()
    with torch.no_grad():
        asr.decoder.char_trans.bias[1] += eos_bias
    all_lens = [64, 120, 92, 148]
    lens = [all_lens[i] for i in ids]
    feat, fl = synth.padded_batch(ids, lens)
    dec = BeamDecoder(asr, None, 1, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=lm_w, ctc_weight=ctc_w).to(cuda)
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    for k, n in enumerate(lens):
        ora = _oracle_nbest(asr, lm, feat[k], n, 1, lm_w, ctc_w)
        assert [len(h.outIndex) for h in out[k]] == [len(o[0]) for o in ora], (k, [h.outIndex for h in out[k]], [o[0].tolist() for o in ora])
        if ora:
            s, t = _compare(out[k], ora, "beam 1 utt %d" % ids[k])
            assert s == 1
    if eos_bias > 0:
        assert all(len(o) == 0 for o in out)        # the fixture is the empty-N-best case


@pytest.mark.parametrize("mode,v_proj,variant", [("dot", False, ""), ("loc", True, ""), ("loc", False, "yaml_encoder"), ("loc", False, "gru")])
def test_decode_other_model_configurations(cuda, mode, v_proj, variant):
    """The reference's other model settings decode like the oracle (pinned to the reference for these settings in
    tests/test_oracle_vs_reference.py): attention.mode = 'dot' (scaled-dot energies, plain PyTorch in the batched stepper — an
    extension: the reference's own beam search raises for this mode, asr.py:331), attention.v_proj (projected values in the fused location-aware kernel), the encoder of
    config/librispeech_asr.yaml as written (no VGG, a layer that drops every other frame: the unpacked encoder path), GRU
    encoder layers and a 2-layer GRU speller (the plain PyTorch cells of the stepper)."""
    import copy
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    cfg = copy.deepcopy(synth.TINY_ASR_CFG)
    cfg["attention"]["mode"], cfg["attention"]["v_proj"] = mode, v_proj
    if variant == "yaml_encoder":
        cfg["encoder"].update({"vgg": 0, "sample_rate": [1, 2], "sample_style": "drop"})
    elif variant == "gru":
        cfg["encoder"]["module"] = "GRU"
        cfg["decoder"].update({"module": "GRU", "layer": 2})
    asr = synth.build_asr(31, cfg, seed=0, peak=4.0)
    _, lm, lm_path, lm_cfg = _models()
    lens = [64, 120, 92, 148]
    feat, fl = synth.padded_batch(list(range(len(lens))), lens)
    dec = BeamDecoder(asr, None, 4, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=0.3, ctc_weight=0.5).to(cuda)
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    same = ties = 0
    for k, n in enumerate(lens):
        ora = _oracle_nbest(asr, lm, feat[k], n, 4, 0.3, 0.5)
        assert len(out[k]) == len(ora)
        s, t = _compare(out[k], ora, "attention %s v_proj %s %s utt %d" % (mode, v_proj, variant, k))
        same, ties = same + s, ties + t
    print("attention %s, v_proj %s, %s: identical 1-best %d/%d, ties %d" % (mode, v_proj, variant or "default", same, len(lens), ties))
    assert same >= len(lens) - 1


class _COracleScorer:
    """The prefix scorer interface of oracle/ctc_prefix_oracle.py on top of the plain-C oracle (bit-identical, tests/
    test_oracle_golden.py; its frame loop is compiled, which the long-form case needs)."""

    def __init__(self, x):
        from oracle import c_oracle as CO
        self.CO, self.x = CO, np.ascontiguousarray(x[0], dtype=np.float32)
        self.input_length, self.odim = self.x.shape

    def init_state(self):
        return self.CO.blank_state(self.x)

    def cheap_compute(self, g, r_prev, candidates):
        return self.CO.extend(self.x, len(g), g[-1] if g else 0, r_prev, candidates)


def _oracle_nbest_c(asr, lm, feat, n, beam, lm_w, ctc_w, max_ratio):
    from oracle import beam_oracle as BO
    asr, lm = _cpu_copy(asr), _cpu_copy(lm)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)                 # batch-1 modules of a few dozen units: intra-op threading only costs
    try:
        with torch.no_grad():
            nb = BO.decode_utterance(asr, feat[None, :n], torch.LongTensor([n]), beam, 0.01, max_ratio, lm=lm if lm_w > 0 else None,
                                     lm_weight=lm_w, ctc_weight=ctc_w, scorer_cls=_COracleScorer)
    finally:
        torch.set_num_threads(threads)
    return BO.nbest_as_arrays(nb)


def test_decode_subword_vocab_10000_matches_oracle(cuda):
    """BASELINE cfg3 end to end at its vocabulary size: V = 10000 (column-gather variant of the fused prefix kernel, the
    one-CTA-per-row posterior kernel, 10k-wide candidate / combine kernels and RNNLM output), beam 8 + LM,
    max_len_ratio 0.07 (config/libri/decode_example.yaml:13), against the per-hypothesis CPU oracle."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, lm, lm_path, lm_cfg = _models(vocab=10000)
    lens = [200, 148, 120]
    feat, fl = synth.padded_batch(list(range(len(lens))), lens)
    dec = BeamDecoder(asr, None, 8, 0.01, 0.07, lm_path=lm_path, lm_config=lm_cfg, lm_weight=0.5, ctc_weight=0.5).to(cuda)
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    same = ties = 0
    for k, n in enumerate(lens):
        ora = _oracle_nbest_c(asr, lm, feat[k], n, 8, 0.5, 0.5, 0.07)
        assert len(out[k]) == len(ora)
        s, t = _compare(out[k], ora, "V 10000 utt %d" % k)
        same, ties = same + s, ties + t
    print("V 10000 beam 8: identical 1-best %d/%d, ties %d" % (same, len(lens), ties))
    assert same + ties == len(lens)


def test_decode_longform_875_frames_beam16_matches_oracle(cuda):
    """BASELINE cfg4 end to end at its sizes: a 35-s utterance (3500 input frames, 875 encoder frames, 700 decode steps),
    beam 16 (24 CTC candidates: 2 state + 2 helper + 12 psi warps per utterance in the fused prefix kernel) + LM, decoded
    together with a shorter one, against the per-hypothesis CPU oracle (small modules so that it finishes in seconds)."""
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    asr, lm, lm_path, lm_cfg = _models()
    lens = [3500, 400]
    feat, fl = synth.padded_batch([40, 41], lens)
    dec = BeamDecoder(asr, None, 16, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=0.3, ctc_weight=0.5).to(cuda)
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    same = ties = 0
    for k, n in enumerate(lens):
        ora = _oracle_nbest_c(asr, lm, feat[k], n, 16, 0.3, 0.5, 0.2)
        assert len(out[k]) == len(ora)
        s, t = _compare(out[k], ora, "long form utt %d (%d frames)" % (k, n))
        same, ties = same + s, ties + t
    print("long form (875 encoder frames, beam 16): identical 1-best %d/%d, ties %d" % (same, len(lens), ties))
    assert same + ties == len(lens)
