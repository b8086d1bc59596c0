"""Sweep E2E_PREFIX_SPLIT_BELOW (the launch size below which kernel (2) uses one-warp CTAs) on the bench workload."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for t in sys.argv[1:]:
    env = dict(os.environ, E2E_PREFIX_SPLIT_BELOW=t)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                         capture_output=True, text=True, env=env).stdout.strip().split("\n")[-1]
    d = json.loads(out)
    print(json.dumps({"split_below": int(t), "utts_per_s": d["value"], "frac": d["roofline"]["frac"],
                      "kernel_ms_per_pass": d["roofline"]["kernel_ms_per_step"]}), flush=True)
