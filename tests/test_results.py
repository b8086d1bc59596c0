"""f-3: result files + scoring (e2e_asr_pytorch_b200.results) against the reference's own writer format
(bin/test_asr.py:146-156) and its eval.py / eval_beam.py scripts, which are RUN here when /root/reference is
present (with a stand-in for the missing `editdistance` package)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from e2e_asr_pytorch_b200 import results as R

REF = "/root/reference"


class CharTok:
    """decode() semantics of CharacterTextEncoder (src/text.py:56-66): stop at <eos>=1, skip <pad>=0."""
    vocab = ["<pad>", "<eos>", "<unk>"] + list(" abcdefghijklmnopqrstuvwxyz'")

    def decode(self, ids):
        out = []
        for i in ids:
            if i == 1:
                break
            if i == 0:
                continue
            out.append(self.vocab[i])
        return "".join(out)


def _dp(a, b):
    d = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        p, d[0] = d[0], i
        for j, y in enumerate(b, 1):
            p, d[j] = d[j], min(d[j] + 1, d[j - 1] + 1, p + (x != y))
    return d[-1]


def test_edit_distance_matches_textbook_dp():
    rng = np.random.default_rng(0)
    for _ in range(300):
        a = "".join(rng.choice(list("abc d"), rng.integers(0, 14)))
        b = "".join(rng.choice(list("abc d"), rng.integers(0, 14)))
        assert R.edit_distance(a, b) == _dp(a, b)
        assert R.edit_distance(a.split(" "), b.split(" ")) == _dp(a.split(" "), b.split(" "))
    assert R.edit_distance("", "abc") == 3 and R.edit_distance("kitten", "sitting") == 3


def _make_files(tmp_path, n_utts=23, beam=4, seed=1):
    rng = np.random.default_rng(seed)
    tok = CharTok()
    results = []
    for u in range(n_utts):
        truth = [int(t) for t in rng.integers(3, 31, rng.integers(5, 40))] + [1]
        hyps = []
        for _ in range(beam):
            h = [t if rng.random() > 0.2 else int(rng.integers(3, 31)) for t in truth[:-1]]
            if rng.random() < 0.3:
                h = h[:max(1, len(h) - 2)]
            hyps.append(h + [1, 5, 6])                      # tokens after <eos> must be cut by decode()
        results.append(("utt-%03d" % u, hyps, truth))
    best, beamf = str(tmp_path / "dev_output.csv"), str(tmp_path / "dev_beam.csv")
    R.init_result_files(best, beamf)
    R.write_results(results, tok, best, beamf)
    return results, tok, best, beamf


def test_result_files_have_the_reference_layout(tmp_path):
    results, tok, best, beamf = _make_files(tmp_path)
    lines = open(best).read().split("\n")
    assert lines[0] == "idx\thyp\ttruth" and len(lines) == len(results) + 2 and lines[-1] == ""
    name, hyps, truth = results[0]
    assert lines[1] == "\t".join([name, tok.decode(hyps[0]), tok.decode(truth)])
    blines = open(beamf).read().split("\n")
    assert blines[0] == "idx\tbeam\thyp\ttruth" and len(blines) == len(results) * 4 + 2
    assert blines[2] == "\t".join([name, "1", tok.decode(hyps[1]), tok.decode(truth)])
    s1, sb = R.score_file(best), R.score_file(beamf, beam=True)
    assert s1["utterances"] == len(results) and sb["utterances"] == len(results) and sb["rows"] == 4 * len(results)
    assert sb["cer_mean"] <= s1["cer_mean"] + 1e-12          # oracle over the N-best can only be better


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
@pytest.mark.parametrize("script,beam", [("eval.py", False), ("eval_beam.py", True)])
def test_scores_match_the_reference_scripts(tmp_path, script, beam):
    """Run the reference's own eval scripts on our files; the printed tables must be identical."""
    _, _, best, beamf = _make_files(tmp_path)
    stub = tmp_path / "stub"
    stub.mkdir()
    (stub / "editdistance.py").write_text(textwrap.dedent('''
        def eval(a, b):
            d = list(range(len(b) + 1))
            for i, x in enumerate(a, 1):
                p, d[0] = d[0], i
                for j, y in enumerate(b, 1):
                    p, d[j] = d[j], min(d[j] + 1, d[j - 1] + 1, p + (x != y))
            return d[-1]
    '''))
    path = beamf if beam else best
    env = dict(os.environ, PYTHONPATH=str(stub), PYTHONDONTWRITEBYTECODE="1")
    ref = subprocess.run([sys.executable, os.path.join(REF, script), "--file", path], capture_output=True, text=True, env=env)
    assert ref.returncode == 0, ref.stderr
    ours = R.format_report(R.score_file(path, beam=beam))
    assert [l.rstrip() for l in ref.stdout.strip("\n").split("\n")] == [l.rstrip() for l in ours.strip("\n").split("\n")]
