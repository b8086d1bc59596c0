"""Which kernels does a decode pass spend its time in?  Kernel-name totals (CUPTI via
torch.profiler, device time only) for BeamDecoder.decode_batch on the bench workload.

    python tools/profile_kernels.py [--n-utts 2620] [--top 40] [--prefix-steps out.json]

``--prefix-steps`` additionally dumps (step, live utterances, prefix-score kernel microseconds)
so the launch-size dependence of kernel (2) can be read off.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-utts", type=int, default=2620)
    ap.add_argument("--max-utts", type=int, default=4096)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--prefix-steps", default="")
    a = ap.parse_args()
    from torch.profiler import profile, ProfilerActivity
    from e2e_asr_pytorch_b200 import shard
    dev = torch.device("cuda:0")
    dec, _, _ = bench.build_models(dev)
    lengths = bench.workload_lengths(1, a.n_utts)
    batches = shard.make_batches(np.arange(len(lengths)), lengths, a.max_utts, 0)
    feats = [bench.make_features(b, lengths, pin=False) for b in batches]
    feats = [(f.to(dev), l.to(dev)) for f, l in feats]
    small = bench.make_features(list(range(8)), lengths, pin=False)
    dec.decode_batch(small[0].to(dev), small[1].to(dev), return_arrays=True)        # warm-up
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for f, l in feats:
            dec.decode_batch(f, l, return_arrays=True)
        torch.cuda.synchronize()
    tot = {}
    prefix = []
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = ev.name
        d = tot.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ev.device_time
        if "prefix_score_kernel" in name:
            prefix.append(ev.device_time)
    total = sum(v[1] for v in tot.values())
    rows = sorted(tot.items(), key=lambda kv: -kv[1][1])
    print("total device time %.1f ms over %d kernels" % (total / 1e3, sum(v[0] for v in tot.values())))
    print("%-100s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
    for name, (n, us) in rows[:a.top]:
        print("%-100s %8d %12.1f %7.4f" % (name[:100], n, us, us / total))
    if a.prefix_steps:
        max_np = np.ceil(lengths * bench.MAX_RATIO).astype(np.int64)
        live = [int((max_np > s).sum()) for s in range(len(prefix))]
        json.dump({"prefix_us": prefix, "live_utts": live}, open(a.prefix_steps, "w"))


if __name__ == "__main__":
    main()
