"""Recipe that compiles the UNMODIFIED reference into ``oracle/_ref/`` (TEST / BASELINE INFRASTRUCTURE).

The reference (DanielLin94144/E2E-ASR-Pytorch) is Python.  Its decode path — ``src/ctc.py``, ``src/decode.py`` and the
modules they import, plus the caller ``bin/test_asr.py`` — is byte-compiled from the sources where they lie under
``/root/reference`` into sourceless bytecode files under ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to
the GPU box like a built ``.so``).  No reference source text enters the repository; the outputs are CPython bytecode
of this image's interpreter (the bytes of a ``.pyc``, stored with the extension ``.refc`` because snapshot tools drop
``*.pyc``; ``oracle/refload.py`` has the importer), and ``MANIFEST.json`` records the sha256 of every source file they were
made from.

Users (and nobody else): ``oracle/refload.py`` → the ``not gpu`` pin tests, ``bench.py --impl reference`` and
``bench.py``'s ``cpu_baseline`` leg (the CPU arm the B200 path is timed against), and the drop-in test that drives
``bin/test_asr.py::beam_decode`` with the product's ``BeamDecoder`` patched in.  The product package never imports it.

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present (the build container); on the GPU box the
prebuilt files are used as they are.
"""
import hashlib
import importlib.util
import json
import os
import py_compile
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("E2E_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(_HERE, "_ref")
# the decode path and its caller: nothing else is staged
EXT = ".refc"
FILES = ["src/ctc.py", "src/decode.py", "src/asr.py", "src/module.py", "src/lm.py", "src/util.py", "bin/__init__.py", "bin/test_asr.py"]


def source_available():
    return os.path.isfile(os.path.join(REF_SRC, "src", "ctc.py"))


def staged():
    return os.path.isfile(os.path.join(OUT, "MANIFEST.json")) and os.path.isfile(os.path.join(OUT, "src", "decode" + EXT))


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(force=False):
    """Byte-compile FILES into oracle/_ref/ (sourceless).  Returns the manifest."""
    if not source_available():
        raise RuntimeError("reference sources not present at %s" % REF_SRC)
    want = {f: _sha(os.path.join(REF_SRC, f)) for f in FILES}
    man_path = os.path.join(OUT, "MANIFEST.json")
    if not force and staged():
        try:
            have = json.load(open(man_path))
            if have.get("sha256") == want and have.get("magic") == importlib.util.MAGIC_NUMBER.hex() and have.get("ext") == EXT:
                return have
        except Exception:
            pass
    for f in FILES:
        dst = os.path.join(OUT, f[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path shown in tracebacks (there is no source file to open on the GPU box)
        py_compile.compile(os.path.join(REF_SRC, f), cfile=dst, dfile="reference/" + f, doraise=True, optimize=0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    man = {"what": "CPython bytecode of the unmodified reference's decode path (no sources)", "from": REF_SRC,
           "python": sys.version.split()[0], "magic": importlib.util.MAGIC_NUMBER.hex(), "ext": EXT, "sha256": want}
    with open(man_path, "w") as f:
        json.dump(man, f, indent=1)
    return man


if __name__ == "__main__":
    print(json.dumps(stage(force=True), indent=1))
