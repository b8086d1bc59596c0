"""Micro-benchmark of the VGG front end's device path (model.VGGFrontEnd.forward_masked_split: direct first layer,
unfold+split kernel, library tensor-core GEMMs, bias/ReLU/mask(/pool) kernels) on one encoder chunk, in both GEMM
operand formats, A/B on the same weights and features.

    python tools/bench_vgg.py [--utts 128] [--frames 720] [--steps 5]

Prints one JSON line per format (ms per chunk) and the max |difference| of the outputs relative to their scale.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=128, help="utterances per encoder chunk (stepper.encode uses 128)")
    ap.add_argument("--frames", type=int, default=720, help="feature frames of the longest utterance (the workload's mean)")
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    from e2e_asr_pytorch_b200.model import VGGFrontEnd, reference_init_
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    vgg = VGGFrontEnd(160)
    vgg.apply(reference_init_)
    vgg.eval().to(dev)
    g = torch.Generator().manual_seed(1)
    lens = torch.sort(torch.randint(a.frames // 2, a.frames + 1, (a.utts,), generator=g) // 4 * 4, descending=True)[0]
    lens[0] = a.frames // 4 * 4
    feat = torch.randn(a.utts, int(lens[0]), 160, generator=g)
    for i, l in enumerate(lens):
        feat[i, int(l):] = 0
    feat, lens = feat.to(dev), lens.to(dev)
    outs = {}
    with torch.no_grad(), _Fp32Math():
        for fmt in ("bf16x3", "fp16x2"):
            vgg.conv_split_format = fmt
            for _ in range(2):
                vgg.forward_masked_split(feat, lens)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                out, _ = vgg.forward_masked_split(feat, lens)
            e1.record()
            torch.cuda.synchronize()
            outs[fmt] = out
            print(json.dumps({"bench": "vgg_chunk", "format": fmt, "utts": a.utts, "frames": int(lens[0]),
                              "ms_per_chunk": e0.elapsed_time(e1) / a.steps}), flush=True)
    diff = float((outs["bf16x3"] - outs["fp16x2"]).abs().max())
    print(json.dumps({"bench": "vgg_chunk", "max_abs_diff_between_formats": diff, "output_scale": float(outs["bf16x3"].abs().max())}))


if __name__ == "__main__":
    main()
