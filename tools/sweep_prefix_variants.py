"""Run a micro-benchmark against tuning builds of the library (lib/variants/lib_<name>.so, built with
-DE2E_PS_MINBLOCKS / -DE2E_PS_UNROLL / -DE2E_AF_MINBLOCKS ...) and print ms per launch for a few launch shapes.

Build a variant next to the product library (git-ignored; it travels to the GPU box with the snapshot):

    cd e2e-asr-pytorch_b200/csrc && mkdir -p ../lib/variants && nvcc -gencode arch=compute_100a,code=sm_100a -O3 \
        -lineinfo -std=c++17 --expt-extended-lambda -Xcompiler -fPIC -shared -DE2E_PS_MINBLOCKS=4 \
        -o ../lib/variants/lib_mb4.so *.cu

    python tools/sweep_prefix_variants.py              # prefix-score micro-benchmark
    python tools/sweep_prefix_variants.py attention    # attention micro-benchmark
    python tools/sweep_prefix_variants.py lazy         # fused per-step prefix kernel

The library under test is selected with E2E_ASR_B200_LIB (e2e-asr-pytorch_b200/_lib.py)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shapes = [["--utts", "2620"], ["--utts", "2620", "--plen", "60", "--skip-dead", "1"], ["--utts", "64", "--frames", "825"],
          ["--utts", "600", "--frames", "400", "--ragged", "1", "--plen", "100", "--skip-dead", "1"]]
if len(sys.argv) > 1 and sys.argv[1] == "attention":
    tool, shapes = "bench_attention.py", [["--ragged", "1"], ["--utts", "2620", "--frames", "824", "--ragged", "1"], ["--utts", "100", "--frames", "824", "--ragged", "1"]]
elif len(sys.argv) > 1 and sys.argv[1] == "lazy":
    tool, shapes = "bench_prefix.py", [["--utts", "2620", "--lazy", "1", "--poly", "1", "--plen", "2"], ["--utts", "2620", "--lazy", "1", "--poly", "1", "--plen", "60"],
                                       ["--utts", "64", "--frames", "825", "--lazy", "1", "--poly", "1", "--plen", "120"],
                                       ["--utts", "600", "--frames", "400", "--ragged", "1", "--lazy", "1", "--poly", "1", "--plen", "100"]]
else:
    tool = "bench_prefix.py"
for lib in sorted(glob.glob(os.path.join(ROOT, "e2e-asr-pytorch_b200", "lib", "variants", "lib_*.so"))):
    env = dict(os.environ, E2E_ASR_B200_LIB=lib)
    ms = []
    for sh in shapes:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool)] + sh, capture_output=True, text=True, env=env)
        try:
            ms.append(round((lambda d: d.get("ms_mean", d.get("ms")))(json.loads(out.stdout.strip().split("\n")[-1])), 4))
        except Exception:
            ms.append(None)
    print(os.path.basename(lib), ms, flush=True)
