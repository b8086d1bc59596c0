"""C-ABI boundary checks that need no GPU: the shared object builds for sm_100a, loads, and
exports exactly the symbols include/e2e_asr_b200.h declares; argument validation returns the
documented error codes before anything is launched."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "e2e_asr_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(e2e_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from e2e_asr_pytorch_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.e2e_abi_version() == 1
    assert [lib.e2e_padded_vocab(v) for v in (1, 31, 32, 33, 10000)] == [4, 32, 32, 36, 10000]


def test_header_cites_the_reference_interfaces():
    src = open(HEADER).read()
    for cite in ("src/ctc.py:19-27", "src/ctc.py:68-108", "src/ctc.py:29-66", "src/decode.py:94-95",
                 "src/decode.py:134-177", "src/decode.py:180-183"):
        assert cite in src, cite


def test_argument_validation_needs_no_device():
    from e2e_asr_pytorch_b200 import _lib
    lib = _lib.load()
    assert lib.e2e_ctc_log_softmax(None, 1, 1, 31, None, 1, None, 32, None) == -1
    assert b"e2e_ctc_log_softmax" in lib.e2e_last_error()
    assert lib.e2e_ctc_init_state(None, 1, 1, 32, None, None, None) == -1
    assert lib.e2e_beam_candidates(None, 31, 1, 1, 31, 3, None, None, None, None) == -1
    assert lib.e2e_beam_finalize(0, 1, *([None] * 11), 1, *([None] * 5), 1, None) == -1
    # the operand-format entry points: null pointers, geometry and scales are refused before any launch
    fake = ctypes.c_void_p(256)                      # a non-null, aligned "device pointer" that is never dereferenced
    assert lib.e2e_lstm_split_rows(None, 8, None, 1, 8, None, 24, 8, 0, None) == -1
    assert lib.e2e_lstm_split_rows(fake, 8, None, 1, 8, fake, 16, 8, 0, None) == -1             # bf16 needs a pitch of 3K
    assert lib.e2e_lstm_split_rows_f16x2(fake, 8, None, 1, 8, fake, 16, 8, 0, 3.0, None) == -1   # scale not a power of two
    assert b"power of two" in lib.e2e_last_error()
    assert lib.e2e_lstm_split_rows_f16x2(fake, 8, None, 1, 8, fake, 8, 8, 0, 2.0, None) == -1    # fp16 needs a pitch of 2K
    assert lib.e2e_lstm_cell_f16x2(fake, 16, 0.0, fake, None, None, fake, None, 1, 4, fake, fake, None, 0, 0, 0, 1.0, None) == -1
    assert lib.e2e_conv3x3_unfold_split_f16x2(fake, fake, 1, 4, 4, 4, 0, 16, None, fake, None) == -1   # no scale word
    assert lib.e2e_conv_bias_relu_mask_scaled(fake, fake, fake, 1, 4, 4, 4, 0, 16, fake, 0.0, None, None) == -1
    assert b"inv_w_scale" in lib.e2e_last_error()
    assert lib.e2e_conv_bias_relu_mask_pool_scaled(fake, fake, fake, 1, 4, 4, 4, fake, None, 1.0, None, None) == -1
    assert lib.e2e_conv1_direct_amax(fake, 640, fake, fake, fake, 1, 4, 40, 4, 128, fake, None, None) == -1


def test_sass_uses_the_tma_engine_and_no_legacy_tensor_path():
    """The rows variant of the prefix-score kernel stages posterior tiles with cp.async.bulk
    (SASS: UBLKCP + SYNCS mbarrier ops); nothing in the library is a tensor-core contraction."""
    import shutil
    import subprocess
    from e2e_asr_pytorch_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTMALDG" in sass and "UBLKCP" in sass and "SYNCS" in sass and "LDGSTS" in sass
    assert "HMMA" not in sass and "HGMMA" not in sass


def test_product_has_no_cpu_path():
    import torch
    from e2e_asr_pytorch_b200 import CTCPrefixScore, BeamDecoder, ops, _lib, synth
    with pytest.raises(_lib.E2EError):
        CTCPrefixScore(torch.zeros(1, 4, 31))
    with pytest.raises(_lib.E2EError):
        ops.ctc_log_softmax(torch.zeros(1, 4, 31))
    asr = synth.build_asr(31, synth.TINY_ASR_CFG)
    dec = BeamDecoder(asr, None, 2, 0.01, 0.2, ctc_weight=0.5)
    with pytest.raises(_lib.E2EError):
        dec(torch.zeros(1, 64, 160), torch.LongTensor([64]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "e2e-asr-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text.replace("``/root/reference", "").replace("/root/reference/src", "") or f.endswith(".py"), f
