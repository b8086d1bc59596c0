"""Micro-benchmark of the fused location-aware attention step (the command ncu profiles).

    python tools/bench_attention.py [--utts 1024] [--frames 180] [--beam 8] [--ragged 1] [--nb 0] [--kernels 1]

``--kernels 1`` adds the device time of each of the two kernels (CUPTI via torch.profiler) and the context
product's HBM floor (one pass over the value rows of the live frames).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=180)
    ap.add_argument("--beam", type=int, default=8)
    ap.add_argument("--dim", type=int, default=300)
    ap.add_argument("--edim", type=int, default=640)
    ap.add_argument("--ragged", type=int, default=0)
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kernels", type=int, default=0)
    a = ap.parse_args()
    from e2e_asr_pytorch_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    U, T, B, A, E = a.utts, a.frames, a.beam, a.dim, a.edim
    n = U * B
    key = torch.tanh(torch.randn(U, T, A, generator=g)).to(dev)
    value = torch.randn(U, T, E, generator=g).to(dev)
    query = torch.tanh(torch.randn(n, A, generator=g)).to(dev)
    enc_len = (torch.randint(T // 4, T + 1, (U,), generator=g) if a.ragged else torch.full((U,), T)).to(torch.int32)
    enc_len, _ = torch.sort(enc_len, descending=True)
    prev = torch.softmax(torch.randn(n, T, generator=g), -1)
    for i in range(n):
        prev[i, int(enc_len[i // B]):] = 0
    prev, enc_len = prev.to(dev), enc_len.to(dev)
    w_conv = (torch.randn(10, 201, generator=g) * 0.07).to(dev)
    w_proj = (torch.randn(A, 10, generator=g) * 0.3).to(dev)
    w_e = (torch.randn(A, generator=g) / A ** 0.5).to(dev)
    attn = torch.empty(n, T, device=dev)
    ctx = torch.empty(n, E, device=dev)
    key_t = key.transpose(1, 2).contiguous()
    run = lambda: ops.attention_loc_full(key_t, value, query, prev, enc_len, w_conv, w_proj, w_e, 0.1, 0.5, B,
                                         hyps_per_unit=a.nb, attn=attn, ctx=ctx)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    evs = []
    for _ in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(x.elapsed_time(y) for x, y in evs) / len(evs)
    frames = float(enc_len.sum().item()) * B
    per_kernel = {}
    if a.kernels:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(a.steps):
                run()
            torch.cuda.synchronize()
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA and "attention" in ev.name:
                key = "energy_us" if "energy" in ev.name else "softmax_context_us"
                per_kernel[key] = per_kernel.get(key, 0.0) + ev.device_time / a.steps
        per_kernel["context_hbm_floor_us"] = float(enc_len.sum().item()) * E * 4 / 6496.8e9 * 1e6
    print(json.dumps({**per_kernel, "kernel": "attention_loc_full", "utts": U, "frames": T, "beam": B, "ragged": a.ragged, "nb": a.nb, "ms": ms,
                      "hyp_frames": frames, "hyp_frame_channels_per_s": frames * A / (ms * 1e-3),
                      "mufu_bound_frac": frames * A * 4 / (ms * 1e-3) / (148 * 16 * 1.9e9)}))


if __name__ == "__main__":
    main()
