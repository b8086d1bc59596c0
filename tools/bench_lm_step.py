"""Micro-benchmark of one RNNLM step of the batched decode (stepper._FusedLstm: split kernel + library tensor-core
GEMMs + cell kernel per layer) in both GEMM operand formats, A/B on the same weights and states.

    python tools/bench_lm_step.py [--rows 20960] [--dim 1024] [--layers 4] [--vocab 31] [--steps 20]

Prints one JSON line per format: ms per step, the tensor-core rate of the partial products actually issued, and the
max |difference| of the top hidden state between the formats after the timed steps (same tokens, same parents).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20960, help="hypotheses advanced per step (2620 utterances x beam 8)")
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--vocab", type=int, default=31)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    from e2e_asr_pytorch_b200.stepper import _FusedLstm
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    rnn = torch.nn.LSTM(a.dim, a.dim, num_layers=a.layers, batch_first=True).to(dev)
    emb = torch.randn(a.vocab, a.dim, device=dev)
    g = torch.Generator().manual_seed(2)
    toks = [torch.randint(0, a.vocab, (a.rows,), generator=g).to(dev) for _ in range(a.steps + 3)]
    perms = [torch.randint(0, a.rows, (a.rows,), generator=g).to(dev) for _ in range(a.steps + 3)]
    tops = {}
    with torch.no_grad(), _Fp32Math():
        for fmt, products in (("bf16x3", 6), ("fp16x2", 3)):
            lstm = _FusedLstm(rnn, emb, split=fmt)
            lstm.start(a.rows, dev)
            for i in range(3):
                lstm.step(a.rows, tok=toks[i])
                lstm.reorder(perms[i])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(3, 3 + a.steps):
                top = lstm.step(a.rows, tok=toks[i])
                lstm.reorder(perms[i])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            tops[fmt] = top.clone()
            k_total = sum(k + a.dim for k in lstm.k_in)                  # contraction length summed over layers
            flops = 2.0 * a.rows * 4 * a.dim * k_total * products
            print(json.dumps({"bench": "lm_step", "format": fmt, "rows": a.rows, "dim": a.dim, "layers": a.layers, "ms_per_step": ms,
                              "partial_products": products, "tensor_tflops": flops / (ms * 1e-3) / 1e12}), flush=True)
    print(json.dumps({"bench": "lm_step", "max_abs_diff_between_formats": float((tops["bf16x3"] - tops["fp16x2"]).abs().max())}))


if __name__ == "__main__":
    main()
