"""Where does a decode pass spend its time?  CUDA-event time per phase of BeamDecoder.decode_batch
on the bench workload (optionally a subset).  python tools/profile_phases.py [--n-utts 2620]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-utts", type=int, default=2620)
    ap.add_argument("--max-utts", type=int, default=4096)
    ap.add_argument("--max-padded-frames", type=int, default=0)
    a = ap.parse_args()
    from e2e_asr_pytorch_b200 import shard
    dev = torch.device("cuda:0")
    dec, _, _ = bench.build_models(dev)
    lengths = bench.workload_lengths(1, a.n_utts)
    batches = shard.make_batches(np.arange(len(lengths)), lengths, a.max_utts, a.max_padded_frames)
    feats = [bench.make_features(b, lengths, pin=False) for b in batches]
    feats = [(f.to(dev), l.to(dev)) for f, l in feats]
    for f, l in feats[-1:]:
        dec.decode_batch(f, l, return_arrays=True)          # warm-up on the shortest batch
    dec.profile_phases = True
    dec.phase_ms = {}
    torch.cuda.synchronize()
    t0 = time.time()
    info = []
    for b, (f, l) in zip(batches, feats):
        t1 = time.time()
        dec.decode_batch(f, l, return_arrays=True)
        torch.cuda.synchronize()
        info.append({"utts": len(b), "Lmax": int(lengths[b[0]]), "Lmin": int(lengths[b[-1]]), "steps": dec.last_stats["steps"],
                     "wall_s": round(time.time() - t1, 3)})
    wall = time.time() - t0
    tot = sum(dec.phase_ms.values())
    print(json.dumps({"wall_s": wall, "utts_per_s": len(lengths) / wall, "gpu_ms_total": tot,
                      "phase_ms": {k: round(v, 1) for k, v in dec.phase_ms.items()},
                      "phase_share": {k: round(v / tot, 4) for k, v in dec.phase_ms.items()}, "batches": info}))


if __name__ == "__main__":
    main()
