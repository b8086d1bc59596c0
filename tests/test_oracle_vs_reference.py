"""Live pin: the restatements under oracle/ against the UNMODIFIED reference imported from
/root/reference.  Only possible in the build container; skipped where the tree is absent
(the GPU box) — tests/golden carries the same evidence there."""
import copy

import numpy as np
import pytest
import torch

from oracle import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


def test_prefix_oracle_bit_exact_vs_reference_class():
    from oracle import ctc_prefix_oracle as O, c_oracle as CO
    from tests._util import posteriors
    ref = refload.load()
    rng = np.random.default_rng(0)
    for (t_len, vocab, n_cand) in [(1, 31, 3), (2, 31, 3), (7, 31, 12), (60, 31, 12), (50, 200, 12), (250, 31, 12)]:
        post = posteriors(rng, 1, t_len, vocab)
        R, P = ref.CTCPrefixScore(torch.from_numpy(post)), O.PrefixScorerOracle(post)
        r = R.init_state()
        assert np.array_equal(r, P.init_state()) and np.array_equal(r, CO.blank_state(post[0]))
        g = []
        for step in range(min(t_len + 2, 14)):
            cs = [int(c) for c in rng.permutation(vocab)[:n_cand]]
            if step % 2 == 0 and 1 not in cs:
                cs[0] = 1
            if g and step % 3 == 0 and g[-1] not in cs:
                cs[-1] = g[-1]
            try:
                a_psi, a_r = R.cheap_compute(g, r, cs)
            except IndexError:
                with pytest.raises(IndexError):
                    P.cheap_compute(g, r, cs)
                break
            b_psi, b_r = P.cheap_compute(g, r, cs)
            c_psi, c_r = CO.extend(post[0], len(g), g[-1] if g else 0, r, cs)
            assert np.array_equal(a_psi, b_psi) and np.array_equal(np.ascontiguousarray(a_r), b_r)
            assert np.array_equal(a_psi, c_psi) and np.array_equal(b_r, c_r)
            f_psi, f_r = R.full_compute(g, r)
            h_psi, h_r = P.full_compute(g, r)
            assert np.array_equal(f_psi, h_psi) and np.array_equal(np.ascontiguousarray(f_r), h_r)
            k = int(rng.integers(n_cand))
            g, r = g + [cs[k]], b_r[k]


def test_beam_oracle_exact_vs_reference_decoder(tmp_path):
    import yaml
    from oracle import beam_oracle as BO
    from e2e_asr_pytorch_b200 import synth
    ref = refload.load()
    torch.set_num_threads(4)
    mine = synth.build_asr(31, synth.TINY_ASR_CFG, seed=0, peak=4.0)
    rasr = ref.ASR(synth.FEAT_DIM, 31, True, **copy.deepcopy(synth.TINY_ASR_CFG)).eval()
    rasr.load_state_dict(mine.state_dict())                 # same parameter names and shapes
    lm = synth.build_lm(31, synth.TINY_LM_CFG, seed=1)
    torch.save({"model": lm.state_dict()}, str(tmp_path / "lm.pth"))
    yaml.safe_dump({"model": synth.TINY_LM_CFG}, open(str(tmp_path / "lm.yaml"), "w"))
    for beam, lm_w, n in [(2, 0.0, 64), (8, 0.5, 92)]:
        feat, fl = synth.utterance(7, n)[None], torch.LongTensor([n])
        rdec = ref.BeamDecoder(rasr, None, beam, 0.01, 0.2, lm_path=str(tmp_path / "lm.pth"),
                               lm_config=str(tmp_path / "lm.yaml"), lm_weight=lm_w, ctc_weight=0.5)
        with torch.no_grad():
            want = rdec(feat, fl)
            got_ref_modules = BO.decode_utterance(rasr, feat, fl, beam, 0.01, 0.2, lm=rdec.lm if lm_w > 0 else None,
                                                  lm_weight=lm_w, ctc_weight=0.5)
            got_our_modules = BO.decode_utterance(mine, feat, fl, beam, 0.01, 0.2, lm=lm if lm_w > 0 else None,
                                                  lm_weight=lm_w, ctc_weight=0.5)
        for a, b, c in zip(want, got_ref_modules, got_our_modules):
            assert a.outIndex == b.ids == c.ids
            assert [float(s) for s in a.output_scores] == [float(s) for s in b.scores] == [float(s) for s in c.scores]
            assert float(a.avgScore()) == float(b.mean_score()) == float(c.mean_score())


@pytest.mark.parametrize("variant,beam,lm_w,ctc_w,eos_bias", [
    ("v_proj", 4, 0.3, 0.5, 0.0), ("yaml_encoder", 4, 0.3, 0.5, 0.0), ("gru", 4, 0.3, 0.5, 0.0),
    ("concat_ln", 4, 0.3, 0.0, 0.0), ("", 1, 0.0, 0.0, 0.0), ("", 1, 0.3, 0.0, 4.0), ("", 1, 0.3, 0.5, 0.0), ("", 24, 0.0, 0.0, 0.0)])
def test_beam_oracle_and_product_modules_exact_vs_reference_on_other_configurations(tmp_path, variant, beam, lm_w, ctc_w, eos_bias):
    """The pin of the configuration branches the GPU tests decode (tests/test_gpu_decode.py): with the same weights the
    reference's BeamDecoder, the oracle on the reference's modules and the oracle on model.py's modules give the same N-best
    bit for bit — value projection, the yaml's encoder (no VGG, frame dropping), GRU layers, frame concatenation + layer norm
    (CTC off: with it this fixture is in the reference's crash envelope, decode.py:252), beam 1 (early return, empty N-best when
    <eos> wins before min_len, one CTC candidate), beam 24.  Scaled-dot attention is not here: the reference's own beam search
    cannot run it (see the next test)."""
    import yaml
    from oracle import beam_oracle as BO
    from e2e_asr_pytorch_b200 import synth
    ref = refload.load()
    torch.set_num_threads(1)
    cfg = copy.deepcopy(synth.TINY_ASR_CFG)
    if variant == "v_proj":
        cfg["attention"]["v_proj"] = True
    elif variant == "yaml_encoder":
        cfg["encoder"].update({"vgg": 0, "sample_rate": [1, 2], "sample_style": "drop"})
    elif variant == "gru":
        cfg["encoder"]["module"] = "GRU"
        cfg["decoder"].update({"module": "GRU", "layer": 2})
    elif variant == "concat_ln":
        cfg["encoder"].update({"sample_rate": [2, 1], "sample_style": "concat", "layer_norm": [True, True], "proj": [False, False]})
    mine = synth.build_asr(31, cfg, seed=0, peak=4.0)
    with torch.no_grad():
        mine.decoder.char_trans.bias[1] += eos_bias
    rasr = ref.ASR(synth.FEAT_DIM, 31, True, **copy.deepcopy(cfg)).eval()
    rasr.load_state_dict(mine.state_dict())
    lm = synth.build_lm(31, synth.TINY_LM_CFG, seed=1)
    torch.save({"model": lm.state_dict()}, str(tmp_path / "lm.pth"))
    yaml.safe_dump({"model": synth.TINY_LM_CFG}, open(str(tmp_path / "lm.yaml"), "w"))
    for utt, n in ((0, 64), (1, 120)):
        feat, fl = synth.utterance(utt, n)[None], torch.LongTensor([n])
        rdec = ref.BeamDecoder(rasr, None, beam, 0.01, 0.2, lm_path=str(tmp_path / "lm.pth"),
                               lm_config=str(tmp_path / "lm.yaml"), lm_weight=lm_w, ctc_weight=ctc_w)
        with torch.no_grad():
            want = rdec(feat, fl)
            got_ref_modules = BO.decode_utterance(rasr, feat, fl, beam, 0.01, 0.2, lm=rdec.lm if lm_w > 0 else None,
                                                  lm_weight=lm_w, ctc_weight=ctc_w)
            got_our_modules = BO.decode_utterance(mine, feat, fl, beam, 0.01, 0.2, lm=lm if lm_w > 0 else None,
                                                  lm_weight=lm_w, ctc_weight=ctc_w)
        assert len(want) == len(got_ref_modules) == len(got_our_modules)
        if eos_bias > 0:
            assert len(want) == 0                                   # the empty-N-best fixture
        for a, b, c in zip(want, got_ref_modules, got_our_modules):
            assert a.outIndex == b.ids == c.ids
            assert [float(s) for s in a.output_scores] == [float(s) for s in b.scores] == [float(s) for s in c.scores]
            assert float(a.avgScore()) == float(b.mean_score()) == float(c.mean_score())


def test_reference_beam_search_cannot_run_scaled_dot_attention():
    """attention.mode = 'dot' trains in the reference but its BeamDecoder raises on the first step: ASR.set_state hands the
    previous alignment to BaseAttention.set_mem(), which takes none (asr.py:331, module.py:1095).  The product decodes this
    configuration (tests/test_gpu_decode.py compares it with the oracle on model.py's modules); it is an extension, not a
    parity case — this test documents why no reference N-best exists for it."""
    from e2e_asr_pytorch_b200 import synth
    ref = refload.load()
    cfg = copy.deepcopy(synth.TINY_ASR_CFG)
    cfg["attention"]["mode"] = "dot"
    rasr = ref.ASR(synth.FEAT_DIM, 31, True, **cfg).eval()
    rdec = ref.BeamDecoder(rasr, None, 2, 0.01, 0.2, ctc_weight=0.5)
    with torch.no_grad(), pytest.raises(TypeError, match="set_mem"):
        rdec(synth.utterance(0, 64)[None], torch.LongTensor([64]))


def test_reference_checkpoint_layout_loads_into_product_model():
    """Parameter names/shapes of model.py equal the reference's at the BASELINE dims."""
    from e2e_asr_pytorch_b200 import synth
    ref = refload.load()
    ours = synth.build_asr(31, seed=0)
    theirs = ref.ASR(synth.FEAT_DIM, 31, True, **copy.deepcopy(synth.ASR_MODEL_CFG))
    a, b = ours.state_dict(), theirs.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    lm_a = synth.build_lm(31).state_dict()
    lm_b = ref.RNNLM(31, **copy.deepcopy(synth.LM_MODEL_CFG)).state_dict()
    assert list(lm_a.keys()) == list(lm_b.keys()) and all(lm_a[k].shape == lm_b[k].shape for k in lm_a)


def test_staged_reference_bytecode_equals_the_live_tree():
    """oracle/_ref (the reference's decode path byte-compiled by oracle/ref_stage.py — what travels to the GPU box for
    bench.py's CPU arm and the drop-in test) behaves like the live source tree: same N-best, same scores, through
    bin/test_asr.py::beam_decode.  Each variant runs in its own interpreter (one import of `src.*` per process)."""
    import json, os, subprocess, sys
    from oracle import ref_stage
    ref_stage.stage()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = r'''
import sys, json, copy, torch
sys.path.insert(0, %r)
from oracle import refload
from e2e_asr_pytorch_b200 import synth
staged = sys.argv[1] == "staged"
ref = refload.load(staged=staged)
ta = refload.load_test_asr(staged=staged)
mine = synth.build_asr(31, synth.TINY_ASR_CFG, seed=0, peak=4.0)
rasr = ref.ASR(synth.FEAT_DIM, 31, True, **copy.deepcopy(synth.TINY_ASR_CFG)).eval()
rasr.load_state_dict(mine.state_dict())
dec = ref.BeamDecoder(rasr, None, 4, 0.01, 0.2, ctc_weight=0.5)
name, hyps, _ = ta.beam_decode((["u"], synth.utterance(7, 92)[None], torch.LongTensor([92]), torch.zeros(1, 1, dtype=torch.long)), dec, "cpu")
with torch.no_grad():
    sc = [float(h.avgScore()) for h in dec(synth.utterance(7, 92)[None], torch.LongTensor([92]))]
print(json.dumps({"kind": ref.kind, "hyps": hyps, "scores": sc}))
''' % root
    outs = {}
    for kind in ("live", "staged"):
        r = subprocess.run([sys.executable, "-c", prog, kind], capture_output=True, text=True, timeout=300, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[kind] = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert outs["live"]["kind"] == "live" and outs["staged"]["kind"] == "staged"
    assert outs["live"]["hyps"] == outs["staged"]["hyps"] and outs["live"]["scores"] == outs["staged"]["scores"]
