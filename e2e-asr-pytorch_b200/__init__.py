"""B200-native joint CTC/attention(+RNNLM) beam-search decode path.

Drop-in replacements for the reference's ``src.ctc.CTCPrefixScore`` and
``src.decode.BeamDecoder`` / ``Hypothesis`` on top of hand-written sm_100a CUDA
kernels reached through a C ABI (``include/e2e_asr_b200.h``).  No CPU fallback.
"""
from . import _lib
from .ctc import CTCPrefixScore
from .decode import BeamDecoder, Hypothesis, CTC_BEAM_RATIO, LOG_ZERO

__all__ = ["CTCPrefixScore", "BeamDecoder", "Hypothesis", "install", "CTC_BEAM_RATIO", "LOG_ZERO"]


def install():
    """Patch an importable reference checkout so that ``bin/test_asr.py`` picks this path up:
    ``from src.decode import BeamDecoder`` (bin/test_asr.py:10) then resolves to the classes
    above.  Call before importing ``bin.test_asr``; see INTEGRATION.md."""
    import src.ctc
    import src.decode
    src.ctc.CTCPrefixScore = CTCPrefixScore
    src.decode.CTCPrefixScore = CTCPrefixScore
    src.decode.BeamDecoder = BeamDecoder
    src.decode.Hypothesis = Hypothesis
