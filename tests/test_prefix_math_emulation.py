"""CPU checks of the prefix-score kernel's log-add-exp evaluators through their fp32 emulation
(tools/emulate_prefix_math.py): the committed polynomial coefficients are what the generator produces, the
polynomial itself is good to 4e-8, and the whole recursion emulated with the table (default) and with the
degree-8 polynomial (opt-in E2E_PREFIX_POLY_MATH) stays within the GPU tests' tolerance of the reference
arithmetic (tests/_util.prefix_tolerance: 1e-4 absolute or 2 ulp)."""
import os
import re

import numpy as np

from tests._util import prefix_tolerance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_polynomial_is_the_generated_one():
    from tools.gen_softplus_poly import coefficients
    text = open(os.path.join(ROOT, "e2e-asr-pytorch_b200", "csrc", "softplus_poly.inc")).read()
    got = np.array([float(v) for v in re.findall(r"=\s*([-0-9.e+]+)f;", text)], np.float32)
    want = coefficients()
    assert got.shape == want.shape == (9,)
    assert np.array_equal(got, want)
    e = np.linspace(0.0, 1.0, 100001)
    assert np.abs(e * np.polynomial.polynomial.polyval(e, got.astype(np.float64)) - np.log1p(e)).max() < 4e-8


def test_emulated_recursion_stays_within_the_kernel_tolerance():
    import tools.emulate_prefix_math as E
    F = np.float32
    ref = E.run_chain(lambda a, b: np.logaddexp(a.astype(F), b.astype(F)).astype(F), 120, 31, 12, 8, 6, 7)
    live = ref > -1e7
    assert live.any() and np.abs(ref[live]).max() > 100            # magnitudes where one fp32 ulp is already 1e-5
    tol = prefix_tolerance(ref[live])
    for name, sp in (("lut", lambda ad: E.softplus_lut(ad, E.ex2_exact)),
                     ("poly8", E.softplus_poly(8, E.ex2_exact)),
                     ("poly8 with 1-ulp ex2 noise", E.softplus_poly(8, E.ex2_noisy)),
                     ("poly8 pairwise", E.softplus_poly(8, E.ex2_exact, estrin=True))):
        got = E.run_chain(E.logaddexp_with(sp), 120, 31, 12, 8, 6, 7)
        assert np.all(got[~live] < -1e7), name
        err = np.abs(got - ref)[live]
        assert (err <= tol).all(), "%s: max err %.3g" % (name, err.max())
        assert (err == 0).mean() > 0.9, name                      # and most scores are bit-equal
