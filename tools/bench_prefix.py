"""Micro-benchmark of the prefix-score kernel alone (the command ncu profiles).

    python tools/bench_prefix.py [--utts 2620] [--frames 180] [--beam 8] [--steps 6] [--fast 0] [--skip-dead 0]

Builds a beam-search-shaped state on the device (random posteriors, random candidates, every
utterance with B live hypotheses at prefix length `--plen`), launches the kernel `--steps`
times ping-ponging the state buffers and reports candidate-frames/s and the algorithmic
HBM GB/s (12 + 12/C bytes per candidate-frame) from CUDA events.
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
    ap.add_argument("--beam", type=int, default=8)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--plen", type=int, default=1)
    ap.add_argument("--fast", type=int, default=0)
    ap.add_argument("--libm", type=int, default=0)
    ap.add_argument("--poly", type=int, default=0, help="1: MUFU.EX2 + polynomial log-add-exp (Horner), 2: the same, pairwise (Estrin)")
    ap.add_argument("--skip-dead", type=int, default=0)
    ap.add_argument("--ragged", type=int, default=0)
    ap.add_argument("--lazy", type=int, default=0, help="1: the fused per-step kernel with lazy state evaluation (e2e_ctc_prefix_step)")
    a = ap.parse_args()
    from e2e_asr_pytorch_b200 import ops, _lib as L
    dev = torch.device("cuda:0")
    U, T, V, B = a.utts, a.frames, a.vocab, a.beam
    C = int(1.5 * B)
    g = torch.Generator(device="cpu").manual_seed(0)
    logits = torch.randn(U, T, V, generator=g).to(dev)
    if a.ragged:
        enc_len = torch.randint(T // 4, T + 1, (U,), generator=g).to(torch.int32).to(dev)
    else:
        enc_len = torch.full((U,), T, dtype=torch.int32, device=dev)
    x = ops.ctc_log_softmax(logits, enc_len, apply_relu=True)
    r0 = ops.ctc_init_state(x, enc_len)
    n_live = torch.full((U,), B, dtype=torch.int32, device=dev)
    cand = torch.stack([torch.randperm(V, generator=g)[:C] for _ in range(U * B)]).to(torch.int32).to(dev)
    last = torch.randint(2, V, (U * B,), generator=g).to(torch.int32).to(dev)
    plen = torch.full((U * B,), a.plen, dtype=torch.int32, device=dev)
    lane = torch.randint(0, B * C, (U * B,), generator=g).to(torch.int32).to(dev)
    bufs = [torch.empty((U, T, B * C, 2), device=dev) for _ in range(2)]
    psi = torch.empty((U * B, C), device=dev)
    status = torch.zeros(U, dtype=torch.int32, device=dev)
    flags = (L.PREFIX_FAST_MATH if a.fast else 0) | (L.PREFIX_LIBM_MATH if a.libm else 0) | (L.PREFIX_SKIP_DEAD_ROWS if a.skip_dead else 0)
    if a.poly and not (a.fast or a.libm):
        flags |= L.PREFIX_POLY_MATH | (L.PREFIX_POLY_ESTRIN if a.poly == 2 else 0)
    # first launch from the empty prefix fills bufs[0] with valid states
    ops.ctc_prefix_score(x, V, enc_len, r0, torch.zeros_like(lane), last, torch.zeros_like(plen),
                         torch.ones_like(n_live), cand, B, C, 0, psi=psi, r_out=bufs[0], status=status)
    lane0 = torch.randint(0, C, (U * B,), generator=g).to(torch.int32).to(dev)      # only slot 0 lanes are valid after step 0
    ops.ctc_prefix_score(x, V, enc_len, bufs[0], lane0, last, plen, n_live, cand, B, C, 0, psi=psi, r_out=bufs[1], status=status)
    cur = 1
    if a.lazy:
        # state buffers hold the B live hypotheses only; every hypothesis' parent is a random slot of the previous buffer
        bufs = [torch.empty((U, T, B, 2), device=dev) for _ in range(2)]
        pslot = torch.randint(0, B, (U * B,), generator=g).to(torch.int32).to(dev)
        ptok = torch.randint(2, V, (U * B,), generator=g).to(torch.int32).to(dev)
        one = torch.ones_like(plen)
        ops.ctc_prefix_step(x, V, enc_len, r0, torch.zeros_like(pslot), last, None, one, n_live, cand, B, C, 0,
                            psi=psi, r_out=bufs[0], status=status)
        ops.ctc_prefix_step(x, V, enc_len, bufs[0], pslot, last, ptok, one + 1, n_live, cand, B, C, 0,
                            psi=psi, r_out=bufs[1], status=status)
        lflags = flags & ~(L.PREFIX_SKIP_DEAD_ROWS)

    def launch():
        nonlocal cur
        if a.lazy:
            ops.ctc_prefix_step(x, V, enc_len, bufs[cur], pslot, last, ptok, plen, n_live, cand, B, C, lflags,
                                psi=psi, r_out=bufs[1 - cur], status=status)
            cur = 1 - cur
            return
        _launch_eager()

    def _launch_eager():
        nonlocal cur
        ops.ctc_prefix_score(x, V, enc_len, bufs[cur], lane, last, plen, n_live, cand, B, C, flags,
                             psi=psi, r_out=bufs[1 - cur], status=status)
        cur = 1 - cur

    for _ in range(a.warmup):
        launch()
    torch.cuda.synchronize()
    evs = []
    for _ in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = np.array([e0.elapsed_time(e1) for e0, e1 in evs])
    units = float(B * C) * float(enc_len.sum().item())
    bpu = 12.0 + 12.0 / C
    peak = 6496.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = units * bpu / (ms.mean() * 1e-3) / 1e9
    print(json.dumps({"kernel": "prefix_step_lazy" if a.lazy else "prefix_score", "utts": U, "frames": T, "vocab": V, "beam": B, "cand": C, "plen": a.plen,
                      "math": "mufu" if a.fast else ("libm" if a.libm else (["lut", "poly", "poly_estrin"][a.poly])), "skip_dead_rows": a.skip_dead, "ms_mean": float(ms.mean()), "ms_min": float(ms.min()),
                      "cand_frames": units, "cand_frames_per_s": units / (ms.mean() * 1e-3), "bytes_per_cand_frame": bpu,
                      "algorithmic_GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak, "status": int(status.sum().item()),
                      "state_buffer_MB": bufs[0].numel() * 4 / 1e6}))


if __name__ == "__main__":
    main()
