"""CPU emulation (numpy, fp32 with emulated FMA) of the prefix-score recursion of csrc/prefix_score.cu with the
kernel's log-add-exp evaluators, against the reference arithmetic (np.logaddexp in fp32 = src/ctc.py:68-108):

    lut           the default: piecewise-cubic table on [-4, 0] + MUFU.EX2 series tail (common.cuh softplus_lut)
    poly7/8/9     MUFU.EX2 + polynomial e*P(e) (common.cuh softplus_poly; the kernel ships degree 8), Horner
    poly8_estrin  the same polynomial evaluated pairwise
    *_noisy       with +-1 ulp of noise on every ex2 (MUFU.EX2's error bound is 2^-22 relative; plain = correctly rounded)

A beam-search-shaped chain is run per evaluator (every step extends each hypothesis by C random candidates and
continues from one of them with ITS OWN states), and the prefix scores are compared with the reference chain:
max error in absolute terms and in ulps of the value, and the share of bit-equal scores.  This is how the
polynomial's degree was chosen without GPU time; it says nothing about speed.

    python tools/emulate_prefix_math.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F = np.float32
LOGZERO = F(-1e8)
LOG2E = F(-1.4426950408889634)
RNG = np.random.default_rng(123)


def fma(a, b, c):
    """fp32 fused multiply-add: the product is exact in float64, one rounding to fp32."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


def ex2_exact(x):
    return np.exp2(x.astype(np.float64)).astype(F)


def ex2_noisy(x):
    return (np.exp2(x.astype(np.float64)) * (1.0 + RNG.uniform(-2.0 ** -22, 2.0 ** -22, size=x.shape))).astype(F)


def load_table():
    rows = []
    for ln in open(os.path.join(ROOT, "e2e-asr-pytorch_b200", "csrc", "softplus_lut.inc")):
        m = re.findall(r"([-0-9.e+]+)f", ln)
        if len(m) == 4:
            rows.append([float(v) for v in m])
    return np.array(rows, F)


TABLE = load_table()


def softplus_lut(ad, ex2):
    magic = F(12582912.0 + 64.0)
    tc = np.minimum(ad, F(4.0))
    tm = fma(tc, F(-16.0), magic)
    f = fma(tc, F(-16.0), (magic - tm).astype(F))
    c = TABLE[(tm.view(np.uint32) - np.uint32(0x4B400000)).astype(np.int64)]
    g = fma(f, fma(f, fma(f, c[..., 3], c[..., 2]), c[..., 1]), c[..., 0])
    e = ex2((ad * LOG2E).astype(F))
    q = fma(e, fma(e, fma(e, F(-0.25), F(0.333333343)), F(-0.5)), F(1.0))
    return np.where(ad > F(4.0), (e * q).astype(F), g)


def softplus_poly(deg, ex2, estrin=False):
    from tools.gen_softplus_poly import coefficients
    c = coefficients(deg)

    def sp(ad):
        e = ex2((ad * LOG2E).astype(F))
        if not estrin:
            acc = np.full_like(e, c[-1])
            for ci in c[-2::-1]:
                acc = fma(acc, e, ci)
        else:
            assert deg == 8
            e2 = (e * e).astype(F)
            e4 = (e2 * e2).astype(F)
            p01, p23, p45, p67 = fma(c[1], e, c[0]), fma(c[3], e, c[2]), fma(c[5], e, c[4]), fma(c[7], e, c[6])
            lo, hi = fma(p23, e2, p01), fma(p67, e2, p45)
            acc = fma(fma(c[8], e4, hi), e4, lo)
        return (e * acc).astype(F)
    return sp


def logaddexp_with(sp):
    def lae(a, b):
        return (np.maximum(a, b) + sp(np.abs((a - b).astype(F)))).astype(F)
    return lae


def cheap_compute(lae, x, r_prev, cs, plen, last):
    """N hypotheses at once: x [T,V], r_prev [N,T,2], cs [N,C], plen / last [N] -> psi [N,C], r [N,T,C,2]."""
    n, c = cs.shape
    t_len = x.shape[0]
    r = np.full((n, t_len, c, 2), LOGZERO, F)
    start = np.maximum(1, plen)
    xc = x[:, cs]
    r[plen == 0, 0, :, 0] = xc[0][plen == 0]
    psi = r[np.arange(n), start - 1, :, 0].copy()
    sum_prev = lae(r_prev[..., 0], r_prev[..., 1])
    phi = np.repeat(sum_prev[:, :, None], c, 2)
    special = (cs == last[:, None]) & (plen > 0)[:, None]
    phi = np.where(special[:, None, :], r_prev[..., 1][:, :, None], phi)
    for t in range(1, t_len):
        act = (t >= start)[:, None]
        nb, bl = r[:, t - 1, :, 0], r[:, t - 1, :, 1]
        nnb = (lae(nb, phi[:, t - 1]) + xc[t]).astype(F)
        nbl = (lae(bl, nb) + x[t, 0]).astype(F)
        npsi = lae(psi, (phi[:, t - 1] + xc[t]).astype(F))
        r[:, t, :, 0] = np.where(act, nnb, r[:, t, :, 0])
        r[:, t, :, 1] = np.where(act, nbl, r[:, t, :, 1])
        psi = np.where(act, npsi, psi)
    return np.where(cs == 1, sum_prev[:, -1][:, None], psi), r


def run_chain(lae, t_len, vocab, n_cand, steps, n_hyps, seed):
    from tests._util import posteriors
    x = posteriors(np.random.default_rng(seed), 1, t_len, vocab)[0]
    r0 = np.full((t_len, 2), LOGZERO, F)
    acc = F(0)
    for t in range(t_len):
        acc = F(acc + x[t, 0])
        r0[t, 1] = acc
    rng = np.random.default_rng(seed + 1)                # the same candidates / survivors for every evaluator
    r_prev, plen, last = np.repeat(r0[None], n_hyps, 0), np.zeros(n_hyps, np.int64), np.zeros(n_hyps, np.int64)
    out = []
    for _ in range(steps):
        cs = np.stack([rng.choice(np.arange(1, vocab), n_cand, replace=False) for _ in range(n_hyps)])
        psi, r = cheap_compute(lae, x, r_prev, cs, plen, last)
        out.append(psi)
        pick = rng.integers(0, n_cand, n_hyps)
        for i in range(n_hyps):
            while cs[i, pick[i]] == 1:                   # <eos> ends a hypothesis: continue from another candidate
                pick[i] = (pick[i] + 1) % n_cand
        r_prev, last, plen = r[np.arange(n_hyps), :, pick, :], cs[np.arange(n_hyps), pick], plen + 1
    return np.stack(out)


def main():
    evaluators = {"lut": lambda ad: softplus_lut(ad, ex2_exact), "lut_noisy": lambda ad: softplus_lut(ad, ex2_noisy)}
    for deg in (7, 8, 9):
        evaluators["poly%d" % deg] = softplus_poly(deg, ex2_exact)
        evaluators["poly%d_noisy" % deg] = softplus_poly(deg, ex2_noisy)
    evaluators["poly8_estrin"] = softplus_poly(8, ex2_exact, estrin=True)
    evaluators["poly8_estrin_noisy"] = softplus_poly(8, ex2_noisy, estrin=True)
    for t_len, vocab, n_cand, steps, n_hyps in [(180, 31, 12, 30, 16), (875, 31, 24, 12, 8)]:
        ref = run_chain(lambda a, b: np.logaddexp(a.astype(F), b.astype(F)).astype(F), t_len, vocab, n_cand, steps, n_hyps, 7)
        live = ref > -1e7
        ulp = np.spacing(np.abs(ref))[live]
        print("T = %d, %d candidates, %d steps x %d hypotheses, max |psi| = %.0f" % (t_len, n_cand, steps, n_hyps, np.abs(ref[live]).max()))
        for name, sp in evaluators.items():
            err = np.abs(run_chain(logaddexp_with(sp), t_len, vocab, n_cand, steps, n_hyps, 7) - ref)[live]
            print("  %-20s max abs err %.3g   max %.2f ulp   mean %.3f ulp   bit-equal %.1f %%"
                  % (name, err.max(), (err / ulp).max(), (err / ulp).mean(), 100 * (err == 0).mean()))


if __name__ == "__main__":
    main()
