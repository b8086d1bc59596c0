"""bench.py's measurement on the BASELINE.json configurations it does not run by default (they are parity-test
cases there; this tool is for the record, not for the driver):

    python tools/bench_config.py --cfg 3      # subword vocabulary (V = 10000), beam 8 + RNNLM, max_len_ratio 0.07, 2620 utterances
    python tools/bench_config.py --cfg 4      # long form: 256 utterances of 35 s (875 encoder frames), beam 16 + RNNLM

It sets bench.py's workload constants, runs its B200 arm unchanged (same timing rules, same JSON line) and rewrites
the ``config.workload`` text.  cfg3 decodes in memory-budgeted batches (shard.estimate_decode_bytes): its posteriors
alone are 86 GB at the full set.  Extra arguments go to bench.py (e.g. --steps 1 --warmup 1 --no-cpu-baseline).
"""
import argparse
import contextlib
import functools
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, required=True, choices=[3, 4])
    ap.add_argument("--n-utts", type=int, default=0)
    a, rest = ap.parse_known_args()
    from e2e_asr_pytorch_b200 import shard
    bench.GOLDEN_APPLIES = False          # the golden N-best fixtures are of the cfg2 workload
    if a.cfg == 3:
        bench.VOCAB, bench.BEAM, bench.MAX_RATIO = 10000, 8, 0.07
        n_utts = a.n_utts or 2620
        what = "cfg3: subword V=10000 (vocab-gather path), VGG+BLSTM CTC-attention + 4x1024 RNNLM, beam 8, ctc 0.5, lm 0.5, " \
               "max_len_ratio 0.07, %d utts/GPU dev-clean-like lengths in memory-budgeted batches, random init" % n_utts
        bytes_fn = functools.partial(shard.estimate_decode_bytes, vocab=10000, beam=8, n_cand=12)
        plain = shard.make_batches
        shard.make_batches = lambda idx, lengths, max_utts=512, max_padded_frames=None: plain(
            idx, lengths, max_utts, max_padded_frames, int(0.8 * 178e9), bytes_fn)
    else:
        bench.VOCAB, bench.BEAM, bench.MAX_RATIO = 31, 16, 0.2
        n_utts = a.n_utts or 256
        what = "cfg4: long form, %d utts/GPU of 3500 input frames (35 s, 875 encoder frames), char V=31, beam 16 (24 CTC " \
               "candidates), ctc 0.5, lm 0.5, max_len_ratio 0.2, random init" % n_utts
        bench.workload_lengths = lambda world, n: np.full(world * n, 3500, dtype=np.int64)
    sys.argv = [sys.argv[0], "--n-utts", str(n_utts)] + rest
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.main()
    for line in out.getvalue().splitlines():
        if line.startswith("{"):
            d = json.loads(line)
            d["config"]["workload"] = what
            d["metric"] = d["metric"].replace("beam-8", "beam-%d" % bench.BEAM)
            if "roofline" in d:
                d["roofline"]["traffic"] = d["roofline"]["traffic_source"] = None   # the ncu capture is of cfg2
            print(json.dumps(d), flush=True)
        elif line:
            print(line, flush=True)


if __name__ == "__main__":
    main()
