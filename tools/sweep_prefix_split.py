"""Sweep a launch-shape knob of kernel (2) on the bench workload:
    python tools/sweep_prefix_split.py E2E_PREFIX_SPLIT_BELOW 296 700 ...   (launch size below which one-warp CTAs are used)
    python tools/sweep_prefix_split.py E2E_PREFIX_SMALL_TILE_FROM 300 1000 ...   (launch size from which the 16-frame tile is used)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
knob = sys.argv[1]
for t in sys.argv[2:]:
    env = dict(os.environ, **{knob: t})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                         capture_output=True, text=True, env=env).stdout.strip().split("\n")[-1]
    d = json.loads(out)
    print(json.dumps({knob: int(t), "utts_per_s": d["value"], "frac": d["roofline"]["frac"],
                      "kernel_ms_per_pass": d["roofline"]["kernel_ms_per_step"]}), flush=True)
