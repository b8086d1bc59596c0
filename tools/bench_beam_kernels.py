"""Micro-benchmark of kernels (3a) e2e_beam_candidates and (3b) e2e_beam_combine_prune on a beam-search-shaped state.

    python tools/bench_beam_kernels.py [--utts 2620] [--vocab 31] [--beam 8]

Reports the algorithmic HBM GB/s (per hypothesis-step: 4V + 8 + 4C bytes for (3a), 8V + 8C + 16B for (3b)) and
the fraction of the measured copy bandwidth.  Both kernels move a few hundred bytes per hypothesis and do a top-k selection
over them, so they are bound by instruction issue / latency, not by HBM — the fraction is reported for the record.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=2620)
    ap.add_argument("--vocab", type=int, default=31)
    ap.add_argument("--beam", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    from e2e_asr_pytorch_b200 import ops
    dev = torch.device("cuda:0")
    U, V, B = a.utts, a.vocab, a.beam
    C = int(1.5 * B)
    n = U * B
    att = torch.randn(n, V, device=dev) * 3
    lm = torch.randn(n, V, device=dev) * 2
    S = 64
    buf = ops.BeamBuffers(U, B, C, S, torch.zeros(U, dtype=torch.int32), torch.full((U,), S, dtype=torch.int32), dev)
    buf.n_live.fill_(B); buf.n_active.fill_(B)
    buf.prefix_len.fill_(5)
    buf.psi.copy_(-torch.rand(n, C, device=dev) * 8 - 1)
    peak = 6496.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {}
    for name in ("candidates", "combine_prune"):
        ms = []
        for it in range(a.steps + 3):
            buf.n_live.fill_(B); buf.n_active.fill_(B); buf.prefix_len.fill_(5)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if name == "candidates":
                ops.beam_candidates(att, U, B, V, C, buf.n_active, buf.att_stats, buf.cand)
            else:
                ops.beam_combine_prune(buf, att, lm, V, 5, 0.5, 0.5, 1.5)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ms.append(e0.elapsed_time(e1))
        ms = np.array(ms)
        bph = (4.0 * V + 8 + 4 * C) if name == "candidates" else (8.0 * V + 8 * C + 16 * B)
        gbs = n * bph / (ms.mean() * 1e-3) / 1e9
        out[name] = {"ms_mean": float(ms.mean()), "ms_min": float(ms.min()), "bytes_per_hyp_step": bph, "algorithmic_GBps": gbs,
                     "frac_of_measured_hbm_peak": gbs / peak}
    print(json.dumps({"kernel": "beam_kernels", "utts": U, "vocab": V, "beam": B, "cand": C, **out}))


if __name__ == "__main__":
    main()
