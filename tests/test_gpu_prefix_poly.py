"""The table-free log-add-exp of the prefix-score kernel (E2E_PREFIX_POLY_MATH: MUFU.EX2 + degree-8 polynomial,
csrc/common.cuh softplus_poly; 52 instead of 67.5 instructions per candidate-frame in the hot loop) against the
oracle, on the same beam-search-shaped chains and to the same tolerance as the default table math
(test_gpu_kernels.py: 1e-4 absolute or 2 ulp).  tools/emulate_prefix_math.py predicts <= 1 ulp.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu      # confirmed on a B200 (profiles/r02_a_pytest_gpu.txt)


def _flags(mode):
    from e2e_asr_pytorch_b200 import _lib as L
    return L.PREFIX_POLY_MATH | (L.PREFIX_POLY_ESTRIN if mode == "estrin" else 0)


@pytest.mark.parametrize("mode", ["horner", "estrin"])
def test_poly_math_cfg2_shape(cuda, mode):
    from tests.test_gpu_kernels import _chain
    rng = np.random.default_rng(12)
    worst = _chain(cuda, rng, n_utts=5, t_lens=[180, 37, 96, 64, 181], vocab=31, beam=8, n_cand=12, n_steps=10, flags=_flags(mode))
    print("poly/%s cfg2 shape: max |gpu - oracle| = %.3g" % (mode, worst))


@pytest.mark.parametrize("mode", ["horner", "estrin"])
def test_poly_math_longform_and_machine_filling(cuda, mode):
    from tests.test_gpu_kernels import _chain
    from e2e_asr_pytorch_b200 import _lib as L
    rng = np.random.default_rng(13)
    worst = _chain(cuda, rng, n_utts=2, t_lens=[875, 640], vocab=31, beam=16, n_cand=24, n_steps=4, flags=_flags(mode))
    print("poly/%s long form: max |gpu - oracle| = %.3g" % (mode, worst))
    rng = np.random.default_rng(14)
    worst = _chain(cuda, rng, n_utts=1100, t_lens=[24] * 1100, vocab=31, beam=8, n_cand=12, n_steps=3,
                   flags=_flags(mode) | L.PREFIX_SKIP_DEAD_ROWS, check_states=False)
    print("poly/%s 1100 utterances (16-frame tiles): max |gpu - oracle| = %.3g" % (mode, worst))


def test_poly_math_large_vocab_gather(cuda):
    from tests.test_gpu_kernels import _chain
    rng = np.random.default_rng(15)
    worst = _chain(cuda, rng, n_utts=2, t_lens=[60, 33], vocab=10000, beam=8, n_cand=12, n_steps=4, flags=_flags("horner"))
    print("poly V=10000: max |gpu - oracle| = %.3g" % worst)
