"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

* ``ctc_prefix_chains.npz``  — outputs of the reference ``CTCPrefixScore`` (src/ctc.py) on seeded
  posteriors: ``init_state`` and chains of ``cheap_compute`` / ``full_compute`` calls, including
  the <eos>-override, repeated-last-token and len(g)==T edge cases.
* ``beam_nbest_tiny.npz``    — N-best (tokens, per-token scores, mean score) of the reference
  ``BeamDecoder`` (src/decode.py) driving the reference ``ASR``/``RNNLM`` loaded with the weights
  of ``e2e_asr_pytorch_b200.synth``'s tiny random-init models, on seeded synthetic utterances.

* ``beam_nbest_fullsize.npz`` — the same for the FULL-SIZE bench models (synth.ASR_MODEL_CFG: librispeech_asr.yaml dims
  with the VGG front end, 19 M parameters; 4x1024 RNNLM; output layers unscaled, exactly what bench.py builds) at the
  bench's decode settings (beam 8, ctc 0.5, lm 0.5, ratios 0.01 / 0.2) on five short utterances and on 16 utterances of the
  bench's own 2620-utterance set at the quantile midpoints of its length distribution (bench.py checks its N-best of these).

The reference cannot travel to the GPU box, these vectors can.
"""
import copy
import os
import sys
import tempfile

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload                      # noqa: E402
from e2e_asr_pytorch_b200 import synth          # noqa: E402
from tests._util import posteriors              # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def prefix_chains(ref):
    rng = np.random.default_rng(2024)
    out, n_chain = {}, 0
    for (t_len, vocab, n_cand, steps) in [(1, 31, 3, 3), (2, 31, 3, 4), (12, 31, 12, 14), (60, 31, 12, 10),
                                          (250, 31, 3, 8), (40, 200, 12, 6), (180, 31, 12, 16)]:
        post = posteriors(rng, 1, t_len, vocab)
        sc = ref.CTCPrefixScore(torch.from_numpy(post))
        r = sc.init_state()
        key = "chain%d" % n_chain
        out[key + "_x"] = post[0]
        out[key + "_r0"] = r
        g, n_ok = [], 0
        for s in range(steps):
            cs = [int(c) for c in rng.permutation(vocab)[:n_cand]]
            if s % 2 == 0 and 1 not in cs:
                cs[0] = 1
            if g and s % 3 == 0 and g[-1] not in cs:
                cs[-1] = g[-1]
            try:
                psi, rr = sc.cheap_compute(g, r, cs)
            except IndexError:
                break
            fpsi, frr = sc.full_compute(g, r)
            k = "%s_s%d" % (key, s)
            out[k + "_g"] = np.array(g, np.int32)
            out[k + "_cand"] = np.array(cs, np.int32)
            out[k + "_rprev"] = np.ascontiguousarray(r)
            out[k + "_psi"] = np.array(psi, np.float32)
            out[k + "_r"] = np.ascontiguousarray(rr)
            out[k + "_fpsi"] = np.array(fpsi, np.float32)
            if vocab <= 31 and t_len <= 60:
                out[k + "_fr"] = np.ascontiguousarray(frr)
            pick = int(rng.integers(n_cand))
            g = g + [cs[pick]]
            r = np.ascontiguousarray(rr[pick])
            n_ok += 1
        out[key + "_steps"] = np.int32(n_ok)
        n_chain += 1
    out["n_chains"] = np.int32(n_chain)
    np.savez_compressed(os.path.join(OUT, "ctc_prefix_chains.npz"), **out)
    print("ctc_prefix_chains.npz: %d chains" % n_chain)


def beam_nbest(ref):
    vocab = 31
    mine = synth.build_asr(vocab, synth.TINY_ASR_CFG, seed=0, peak=4.0)
    rasr = ref.ASR(synth.FEAT_DIM, vocab, True, **copy.deepcopy(synth.TINY_ASR_CFG)).eval()
    rasr.load_state_dict(mine.state_dict())
    lm = synth.build_lm(vocab, synth.TINY_LM_CFG, seed=1)
    tmp = tempfile.mkdtemp()
    torch.save({"model": lm.state_dict()}, os.path.join(tmp, "lm.pth"))
    yaml.safe_dump({"model": synth.TINY_LM_CFG}, open(os.path.join(tmp, "lm.yaml"), "w"))
    out, n_case = {}, 0
    for beam, lm_w, utt, n in [(2, 0.0, 0, 64), (2, 0.0, 1, 120), (8, 0.5, 1, 120), (8, 0.5, 2, 92),
                               (4, 0.3, 3, 200), (8, 0.5, 4, 76), (8, 0.5, 5, 148), (16, 0.5, 6, 100)]:
        dec = ref.BeamDecoder(rasr, None, beam, 0.01, 0.2, lm_path=os.path.join(tmp, "lm.pth"),
                              lm_config=os.path.join(tmp, "lm.yaml"), lm_weight=lm_w, ctc_weight=0.5)
        feat = synth.utterance(utt, n)[None]
        with torch.no_grad():
            hyps = dec(feat, torch.LongTensor([n]))
        k = "case%d" % n_case
        out[k + "_beam"], out[k + "_lm_w"], out[k + "_utt"], out[k + "_len"] = np.int32(beam), np.float32(lm_w), np.int32(utt), np.int32(n)
        out[k + "_nbest"] = np.int32(len(hyps))
        for j, h in enumerate(hyps):
            out["%s_tok%d" % (k, j)] = np.array(h.outIndex, np.int32)
            out["%s_sc%d" % (k, j)] = np.array([float(s) for s in h.output_scores], np.float32)
            out["%s_avg%d" % (k, j)] = np.float32(float(h.avgScore()))
        n_case += 1
    out["n_cases"] = np.int32(n_case)
    np.savez_compressed(os.path.join(OUT, "beam_nbest_tiny.npz"), **out)
    print("beam_nbest_tiny.npz: %d cases" % n_case)


FULLSIZE_SHORT = [(900001, 152), (900002, 240), (900003, 320), (900004, 480), (900005, 640)]      # (utterance id, input frames)


def fullsize_cases(n_set=16):
    """The five short fixtures + ``n_set`` utterances OF THE BENCH'S OWN SET (cfg2: 2620 dev-clean-like lengths, seed 2) at the
    quantile midpoints of its length distribution — bench.py compares its N-best of exactly these utterances."""
    lengths = synth.devclean_lengths(2620, seed=2)
    order = np.argsort(lengths, kind="stable")
    pos = ((np.arange(n_set) + 0.5) / n_set * len(order)).astype(np.int64)
    return FULLSIZE_SHORT + [(int(order[p]), int(lengths[order[p]])) for p in pos]


_FS = {}


CFG1_CASES = [(910001 + i, 1000) for i in range(6)]        # BASELINE configs[0]: 10-second utterances, beam 2, no LM


def _cfg1_worker(job):
    return _fullsize_worker(job, beam=2, lm_w=0.0, slot="dec_cfg1")


def beam_nbest_cfg1(procs=6):
    """BASELINE configs[0] (the reference's own CPU-runnable case): V = 31, beam 2 (3 CTC candidates), CTC weight 0.5, no
    LM, 1000 input frames, the full-size model — six utterances through the unmodified reference."""
    import multiprocessing as mp
    jobs = [(k, utt, n) for k, (utt, n) in enumerate(CFG1_CASES)]
    with mp.get_context("fork").Pool(procs) as pool:
        res = dict(pool.imap_unordered(_cfg1_worker, jobs))
    out = {"beam": np.int32(2), "lm_w": np.float32(0.0), "ctc_w": np.float32(0.5), "n_cases": np.int32(len(jobs))}
    for k, (utt, n) in enumerate(CFG1_CASES):
        hyps = res[k]
        key = "case%d" % k
        out[key + "_utt"], out[key + "_len"], out[key + "_nbest"] = np.int32(utt), np.int32(n), np.int32(len(hyps))
        for j, (tok, sc, avg) in enumerate(hyps):
            out["%s_tok%d" % (key, j)], out["%s_sc%d" % (key, j)], out["%s_avg%d" % (key, j)] = tok, sc, avg
        print("  cfg1 case %d: utt %d, %d frames, %d tokens, best mean score %.6f, runner-up gap %.3g"
              % (k, utt, n, len(hyps[0][0]), float(hyps[0][2]), float(hyps[0][2]) - float(hyps[1][2])), flush=True)
    np.savez_compressed(os.path.join(OUT, "beam_nbest_cfg1.npz"), **out)
    print("beam_nbest_cfg1.npz: %d cases" % len(jobs))


def _fullsize_worker(job, beam=8, lm_w=0.5, slot="dec"):
    k, utt, n = job
    torch.set_num_threads(1)
    if slot not in _FS:
        ref = refload.load()
        vocab, ctc_w = 31, 0.5
        mine = synth.build_asr(vocab, seed=0)
        rasr = ref.ASR(synth.FEAT_DIM, vocab, True, **copy.deepcopy(synth.ASR_MODEL_CFG)).eval()
        rasr.load_state_dict(mine.state_dict())
        lm = synth.build_lm(vocab, seed=1)
        tmp = tempfile.mkdtemp()
        torch.save({"model": lm.state_dict()}, os.path.join(tmp, "lm.pth"))
        yaml.safe_dump({"model": synth.LM_MODEL_CFG}, open(os.path.join(tmp, "lm.yaml"), "w"))
        _FS[slot] = ref.BeamDecoder(rasr, None, beam, 0.01, 0.2, lm_path=os.path.join(tmp, "lm.pth"),
                                    lm_config=os.path.join(tmp, "lm.yaml"), lm_weight=lm_w, ctc_weight=ctc_w)
    with torch.no_grad():
        hyps = _FS[slot](synth.utterance(utt, n)[None], torch.LongTensor([n]))
    return k, [(np.array(h.outIndex, np.int32), np.array([float(s) for s in h.output_scores], np.float32),
                np.float32(float(h.avgScore()))) for h in hyps]


def beam_nbest_fullsize(ref, procs=8):
    """One utterance per worker process, the reference BeamDecoder with one torch thread each (longest first)."""
    import multiprocessing as mp
    cases = fullsize_cases()
    jobs = sorted([(k, utt, n) for k, (utt, n) in enumerate(cases)], key=lambda j: -j[2])
    with mp.get_context("fork").Pool(procs) as pool:
        res = dict(pool.imap_unordered(_fullsize_worker, jobs))
    out = {"beam": np.int32(8), "lm_w": np.float32(0.5), "ctc_w": np.float32(0.5), "n_cases": np.int32(len(cases))}
    for k, (utt, n) in enumerate(cases):
        hyps = res[k]
        key = "case%d" % k
        out[key + "_utt"], out[key + "_len"], out[key + "_nbest"] = np.int32(utt), np.int32(n), np.int32(len(hyps))
        for j, (tok, sc, avg) in enumerate(hyps):
            out["%s_tok%d" % (key, j)], out["%s_sc%d" % (key, j)], out["%s_avg%d" % (key, j)] = tok, sc, avg
        print("  full-size case %d: utt %d, %d frames, %d tokens, best mean score %.6f, runner-up gap %.3g"
              % (k, utt, n, len(hyps[0][0]), float(hyps[0][2]), float(hyps[0][2]) - float(hyps[1][2])), flush=True)
    np.savez_compressed(os.path.join(OUT, "beam_nbest_fullsize.npz"), **out)
    print("beam_nbest_fullsize.npz: %d cases" % len(cases))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = refload.load()
    torch.set_num_threads(4)
    if len(sys.argv) > 1 and sys.argv[1] == "fullsize":      # leaves the other fixtures untouched
        beam_nbest_fullsize(ref)
    elif len(sys.argv) > 1 and sys.argv[1] == "cfg1":
        beam_nbest_cfg1()
    else:
        prefix_chains(ref)
        beam_nbest(ref)
        beam_nbest_fullsize(ref)
        beam_nbest_cfg1()
