"""Import the UNMODIFIED reference from /root/reference (TEST INFRASTRUCTURE).

Only usable in the build container: ``/root/reference`` does not exist on the
GPU box, so nothing that runs there may call :func:`load`.  It is used by
``tools/make_golden.py`` (fixture generation) and by the ``not gpu`` tests that
pin the restatements in ``oracle/`` against the live reference.

``src/util.py:7-9`` imports matplotlib at module top; matplotlib is absent in
this image, so a stub module is installed first.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("E2E_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "ctc.py"))


def load():
    """Returns a namespace with the reference's ASR, RNNLM, BeamDecoder,
    Hypothesis, CTCPrefixScore classes and the repo root."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True      # the reference tree is read-only
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from src.asr import ASR
    from src.lm import RNNLM
    from src.decode import BeamDecoder, Hypothesis
    from src.ctc import CTCPrefixScore
    ns = types.SimpleNamespace(ASR=ASR, RNNLM=RNNLM, BeamDecoder=BeamDecoder,
                               Hypothesis=Hypothesis, CTCPrefixScore=CTCPrefixScore,
                               root=REFERENCE_ROOT)
    return ns
