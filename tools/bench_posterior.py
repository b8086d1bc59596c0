"""Micro-benchmark of kernel (1), the fused ReLU + log-softmax of the CTC head into the frame-major posterior tensor.

    python tools/bench_posterior.py [--utts 2620] [--frames 180] [--vocab 31] [--ragged 1]

Reports the algorithmic HBM GB/s (8*V + 4 bytes per valid frame-row: logits read once, log-probs written once,
blank running sum) and its fraction of the measured copy bandwidth, from CUDA events.  With --ragged the rows beyond an
utterance's length are padding the kernel fills with log-zero (written, not counted).
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
    ap.add_argument("--frames", type=int, default=180)
    ap.add_argument("--vocab", type=int, default=31)
    ap.add_argument("--ragged", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    from e2e_asr_pytorch_b200 import ops
    dev = torch.device("cuda:0")
    U, T, V = a.utts, a.frames, a.vocab
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(U, T, V, device=dev)
    if a.ragged:
        enc_len = torch.randint(T // 4, T + 1, (U,), generator=g).to(torch.int32).to(dev)
    else:
        enc_len = torch.full((U,), T, dtype=torch.int32, device=dev)
    x = torch.empty((T, U, ops.padded_vocab(V)), device=dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev)          # 256 MB: larger than the 126 MB L2
    for _ in range(3):
        ops.ctc_log_softmax(logits, enc_len, True, out=x)
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.ctc_log_softmax(logits, enc_len, True, out=x); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = np.array(ms)
    rows = float(enc_len.sum().item())
    bytes_alg = rows * (8.0 * V + 4.0)
    peak = 6496.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = bytes_alg / (ms.mean() * 1e-3) / 1e9
    # the empty-prefix state from the posteriors just written (e2e_ctc_init_state): 4 bytes read (one 32-byte sector in HBM terms)
    # and 8 written per valid frame
    r0 = ops.ctc_init_state(x, enc_len)
    torch.cuda.synchronize()
    ms0 = []
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.ctc_init_state(x, enc_len, out=r0); e1.record()
        torch.cuda.synchronize()
        ms0.append(e0.elapsed_time(e1))
    print(json.dumps({"kernel": "ctc_init_state", "utts": U, "frames": T, "vocab": V, "ragged": a.ragged, "ms_mean": float(np.mean(ms0)),
                      "ms_min": float(np.min(ms0)), "frame_rows": rows, "GBps_12_bytes_per_frame": rows * 12.0 / (float(np.mean(ms0)) * 1e-3) / 1e9}))
    print(json.dumps({"kernel": "ctc_log_softmax", "utts": U, "frames": T, "vocab": V, "ragged": a.ragged, "ms_mean": float(ms.mean()),
                      "ms_min": float(ms.min()), "frame_rows": rows, "bytes_per_frame_row": 8.0 * V + 4.0, "algorithmic_GBps": gbs,
                      "frac_of_measured_hbm_peak": gbs / peak, "l2": "flushed between timed launches (256 MB write)"}))


if __name__ == "__main__":
    main()
