"""Experiment: encoder time on the bench workload vs chunk size / cudnn.benchmark (not part of the product)."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from e2e_asr_pytorch_b200.stepper import BatchedStepper
from e2e_asr_pytorch_b200.decode import _Fp32Math

dev = torch.device("cuda:0")
dec, asr, lm = bench.build_models(dev)
lengths = bench.workload_lengths(1, 2620)
order = np.argsort(-lengths, kind="stable")
f, l = bench.make_features(list(order), lengths, pin=False)
f, l = f.to(dev), l.to(dev)
st = BatchedStepper(dec.asr, None, False, False)
ref = None
with torch.no_grad(), _Fp32Math():
    for bm in (False, True):
        torch.backends.cudnn.benchmark = bm
        for chunk in (128, 32, 64, 192):
            for rep in range(2):
                torch.cuda.synchronize(); t0 = time.time()
                enc, el = st.encode(f, l, chunk=chunk)
                torch.cuda.synchronize(); dt = time.time() - t0
            if ref is None: ref = enc.clone()
            m = (torch.arange(enc.shape[1], device=dev)[None, :] < el[:, None])[:, :, None]
            err = ((enc - ref) * m).abs().max().item()
            print(json.dumps({"cudnn_benchmark": bm, "chunk": chunk, "encode_s": round(dt, 3), "max_abs_diff_vs_first": err}), flush=True)
