"""Import the UNMODIFIED reference (TEST / BASELINE INFRASTRUCTURE).

Two places it can come from:

* ``/root/reference`` — the live source tree, present only in the build container.  Used by
  ``tools/make_golden.py`` (fixture generation) and by the ``not gpu`` tests that pin the restatements
  in ``oracle/`` against the reference.
* ``oracle/_ref/`` — the same files byte-compiled by ``oracle/ref_stage.py`` (sourceless bytecode with the
  extension ``.refc``, git-ignored, shipped to the GPU box; imported through the small finder below).  Used there by ``bench.py --impl reference``, ``bench.py``'s
  ``cpu_baseline`` leg and the drop-in test.  TorchScript needs source text, so the two
  ``torch.jit.ScriptModule`` classes of ``src/module.py`` (liGRU, not on the decode path) are defined with
  scripting switched off for the duration of the import; nothing else differs.

``src/util.py:7-9`` imports matplotlib at module top; matplotlib is absent in this image, so a stub
module is installed first.  ``bin/test_asr.py`` imports ``src.solver``
(tensorboard) and ``src.data`` (librosa): stubs stand in for them while it is imported — only its
module-level ``beam_decode`` function (``bin/test_asr.py:159-173``) is used.
"""
import os
import sys
import types

from . import ref_stage

REFERENCE_ROOT = os.environ.get("E2E_REFERENCE_ROOT", "/root/reference")
_loaded = None       # (kind, root)


def available():
    """The live source tree is present (build container only)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "ctc.py"))


def staged_available():
    return ref_stage.staged()


def _stub_matplotlib():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


class _StagedFinder:
    """Meta-path finder / loader for the staged reference: module ``a.b`` <- ``oracle/_ref/a/b.refc`` (the bytes of a
    .pyc: 16-byte header + marshalled code); ``a`` alone is a namespace-like package rooted at ``oracle/_ref/a``."""

    def __init__(self, root):
        self.root = root

    def _path(self, name):
        return os.path.join(self.root, *name.split("."))

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery as M
        if name.split(".")[0] not in ("src", "bin"):
            return None
        base = self._path(name)
        if os.path.isfile(base + ref_stage.EXT):
            return M.ModuleSpec(name, self, origin=base + ref_stage.EXT)
        if os.path.isfile(os.path.join(base, "__init__" + ref_stage.EXT)):
            spec = M.ModuleSpec(name, self, origin=os.path.join(base, "__init__" + ref_stage.EXT), is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        if os.path.isdir(base):
            spec = M.ModuleSpec(name, self, origin=None, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal
        origin = module.__spec__.origin
        if origin is None:
            return                                       # a directory without __init__: nothing to run
        with open(origin, "rb") as f:
            data = f.read()
        exec(marshal.loads(data[16:]), module.__dict__)


class _NoScript:
    """torch.jit scripting off while sourceless modules are imported (TorchScript wants source text)."""

    def __enter__(self):
        import torch.jit._state as st
        self.st, self.prev = st, st._enabled.enabled
        st._enabled.enabled = False

    def __exit__(self, *exc):
        self.st._enabled.enabled = self.prev


def _root(staged):
    global _loaded
    if staged is None:
        staged = not available()
    kind = "staged" if staged else "live"
    if _loaded is not None:
        if _loaded[0] != kind:
            raise RuntimeError("the reference is already imported from the %s tree in this process" % _loaded[0])
        return _loaded
    if staged:
        if not staged_available():
            raise RuntimeError("oracle/_ref is not built: run __graft_entry__.build() where /root/reference exists")
        root = ref_stage.OUT
        if not any(isinstance(f, _StagedFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _StagedFinder(root))
    else:
        if not available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        root = REFERENCE_ROOT
        sys.dont_write_bytecode = True      # the reference tree is read-only
    _stub_matplotlib()
    if not staged and root not in sys.path:
        sys.path.insert(0, root)
    _loaded = (kind, root)
    return _loaded


def load(staged=None):
    """Namespace with the reference's ASR, RNNLM, BeamDecoder, Hypothesis, CTCPrefixScore classes.
    ``staged=None``: the live tree if present, else ``oracle/_ref``; True / False force one of them."""
    kind, root = _root(staged)
    with _NoScript() if kind == "staged" else _Null():
        from src.asr import ASR
        from src.lm import RNNLM
        from src.decode import BeamDecoder, Hypothesis
        from src.ctc import CTCPrefixScore
    return types.SimpleNamespace(ASR=ASR, RNNLM=RNNLM, BeamDecoder=BeamDecoder, Hypothesis=Hypothesis,
                                 CTCPrefixScore=CTCPrefixScore, root=root, kind=kind)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def load_test_asr(staged=None, fresh=False):
    """The reference's ``bin.test_asr`` module (for ``beam_decode``, bin/test_asr.py:159-173).  ``fresh``: import it
    again even if it already is (its ``from src.decode import BeamDecoder`` binds at import time)."""
    kind, root = _root(staged)
    load(staged)
    if fresh:
        sys.modules.pop("bin.test_asr", None)
    added = []
    for name, attrs in (("src.solver", {"BaseSolver": object}), ("src.data", {"load_dataset": None, "load_wav_dataset": None})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            added.append(name)
    try:
        with _NoScript() if kind == "staged" else _Null():
            import bin.test_asr as mod
    finally:
        for name in added:
            sys.modules.pop(name, None)
    return mod


def worker_init(staged=None, threads=None):
    """Initialiser of a joblib/loky worker process: the same import environment as the parent (path, matplotlib /
    solver / data stand-ins), so that a pickled ``partial(beam_decode, model=decoder)`` can be rebuilt there.
    ``threads``: torch intra-op threads of the worker (None = the library default)."""
    load_test_asr(staged)
    if threads is not None:
        import torch
        torch.set_num_threads(int(threads))
