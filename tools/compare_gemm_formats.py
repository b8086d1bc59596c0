"""Decode the bench workload (bench.py, cfg2) once per GEMM operand format and compare the N-best lists:
how many utterances keep the identical 1-best, and for the others whether the two 1-best hypotheses are a
score tie (the fp16x2 winner is in the bf16x3 N-best within TIE_TOL of its best mean score).

    python tools/compare_gemm_formats.py [--n-utts 2620] [--lm fp16x2] [--vgg bf16x3] [--control]

``--control`` measures the workload's own sensitivity instead: the second arm keeps the base formats but decodes the set as
two interleaved half batches, so the library GEMMs see other shapes and round differently in the last bit — the
noise floor any two fp32 implementations of the model step differ by.

Prints one JSON line: seconds per pass of each arm (one warm-up pass each) and the comparison.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

TIE_TOL = 5e-4      # tests/test_gpu_decode.py


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-utts", type=int, default=bench.N_UTTS)
    ap.add_argument("--lm", default="fp16x2")
    ap.add_argument("--vgg", default="bf16x3")
    ap.add_argument("--control", action="store_true")
    a = ap.parse_args()
    if a.control:
        a.lm = a.vgg = "bf16x3"
    dev = torch.device("cuda:0")
    dec, _, _ = bench.build_models(dev)
    lengths = bench.workload_lengths(1, a.n_utts)
    ids = list(np.argsort(-lengths, kind="stable"))
    feat, fl = bench.make_features(ids, lengths, pin=True)
    feat, fl = feat.to(dev), fl.to(dev)
    arms = {}
    for name, lm, vgg in (("base", "bf16x3", "bf16x3"), ("test", a.lm, a.vgg)):
        dec.lm_split, dec.vgg_split = lm, vgg
        halves = name == "test" and a.control

        def one_pass():
            if not halves:
                return [x.numpy() for x in dec.decode_batch(feat, fl, return_arrays=True)]
            parts = [dec.decode_batch(feat[h::2].contiguous(), fl[h::2].contiguous(), return_arrays=True) for h in (0, 1)]
            width = max(p[0].shape[2] for p in parts)
            merged = []
            for j in range(5):
                full = None
                for h in (0, 1):
                    x = parts[h][j]
                    if j < 2:
                        x = torch.nn.functional.pad(x, (0, width - x.shape[2]))
                    if full is None:
                        full = torch.zeros((len(ids),) + tuple(x.shape[1:]), dtype=x.dtype)
                    full[h::2] = x
                merged.append(full.numpy())
            return merged

        one_pass()
        torch.cuda.synchronize()
        t0 = time.time()
        out = one_pass()
        torch.cuda.synchronize()
        arms[name] = (time.time() - t0, out)
    (tb, (tok_b, sc_b, ln_b, avg_b, n_b)), (tt, (tok_t, sc_t, ln_t, avg_t, n_t)) = arms["base"], arms["test"]
    same1 = same_all = ties = unexplained = 0
    worst_gap = 0.0
    for u in range(len(ids)):
        seq = lambda tok, ln, k: tok[u, k, :ln[u, k]].tolist()
        b0, t0_ = seq(tok_b, ln_b, 0), seq(tok_t, ln_t, 0)
        if b0 == t0_:
            same1 += 1
            if n_b[u] == n_t[u] and all(seq(tok_b, ln_b, k) == seq(tok_t, ln_t, k) for k in range(int(n_b[u]))):
                same_all += 1
            continue
        hit = [k for k in range(int(n_b[u])) if seq(tok_b, ln_b, k) == t0_]
        gap = abs(float(avg_b[u, hit[0]]) - float(avg_b[u, 0])) if hit else float("inf")
        if gap < TIE_TOL:
            ties += 1
            worst_gap = max(worst_gap, gap)
        else:
            unexplained += 1
    score_diff = float(np.abs(avg_b[:, 0] - avg_t[:, 0]).max())
    print(json.dumps({"utterances": len(ids), "base": {"lm": "bf16x3", "vgg": "bf16x3", "s_per_pass": tb},
                      "test": {"lm": a.lm, "vgg": a.vgg, "s_per_pass": tt, "two_half_batches": bool(a.control)},
                      "identical_1best": same1, "identical_nbest": same_all, "ties_within_%g" % TIE_TOL: ties,
                      "worst_tie_gap": worst_gap, "unexplained": unexplained, "max_1best_mean_score_diff": score_diff}))


if __name__ == "__main__":
    main()
