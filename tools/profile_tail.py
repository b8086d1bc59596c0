"""Is the tail of a decode bound by the host (launch rate) or by the device?

    python tools/profile_tail.py [--from-step 300]

Decodes the bench workload once (after a warm-up pass) and, for the decode steps from ``--from-step`` on, compares the HOST
time spent issuing them with the DEVICE time they take (CUDA events at the same two points of the stream; the stream is
drained before the first point), then runs the same window again under cProfile and prints where the host time goes.
"""
import argparse
import cProfile
import io
import json
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from-step", type=int, default=300)
    ap.add_argument("--n-utts", type=int, default=2620)
    a = ap.parse_args()
    from e2e_asr_pytorch_b200 import ops
    dev = torch.device("cuda:0")
    dec, _, _ = bench.build_models(dev)
    lengths = bench.workload_lengths(1, a.n_utts)
    ids = list(range(len(lengths)))
    feat, fl = bench.make_features(ids, lengths, pin=False)
    feat, fl = feat.to(dev), fl.to(dev)
    dec.decode_batch(feat, fl, return_arrays="device")          # warm-up
    out = {}
    for mode in ("timed", "cprofile"):
        rec = {}
        prof = cProfile.Profile()
        real = ops.beam_combine_prune
        state = {"on": False}

        def hooked(buf, att, lm, vocab, step, *args, **kw):
            r = real(buf, att, lm, vocab, step, *args, **kw)
            if step == a.from_step - 1:                          # the window starts after this step's last launch
                torch.cuda.synchronize()
                rec["e0"] = torch.cuda.Event(enable_timing=True); rec["e0"].record()
                rec["t0"] = time.perf_counter(); rec["l0"] = ops.launch_count()
                if mode == "cprofile":
                    prof.enable(); state["on"] = True
            rec["last"] = step
            return r

        ops.beam_combine_prune = hooked
        import e2e_asr_pytorch_b200.decode as D
        real_fin = D.ops.beam_finalize

        def fin(buf, *args, **kw):
            if state["on"]:
                prof.disable(); state["on"] = False
            rec["t1"] = time.perf_counter(); rec["l1"] = ops.launch_count()
            rec["e1"] = torch.cuda.Event(enable_timing=True); rec["e1"].record()
            return real_fin(buf, *args, **kw)

        D.ops.beam_finalize = fin
        try:
            dec.decode_batch(feat, fl, return_arrays="device")
        finally:
            ops.beam_combine_prune = real
            D.ops.beam_finalize = real_fin
        torch.cuda.synchronize()
        steps = rec["last"] - a.from_step + 1
        out[mode] = {"steps": steps, "host_issue_ms": (rec["t1"] - rec["t0"]) * 1e3, "device_ms": rec["e0"].elapsed_time(rec["e1"]),
                     "own_launches": rec["l1"] - rec["l0"]}
        if mode == "cprofile":
            s = io.StringIO()
            pstats.Stats(prof, stream=s).sort_stats("tottime").print_stats(22)
            out["cprofile_top"] = s.getvalue().splitlines()[4:40]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
