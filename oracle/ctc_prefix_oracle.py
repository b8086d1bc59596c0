"""numpy restatement of the CTC prefix-score recursion (TEST INFRASTRUCTURE).

Follows ``/root/reference/src/ctc.py`` (``CTCPrefixScore``; Watanabe et al.
TR2017-190, Alg. 2):

* constants ``logzero=-1e8, blank=0, eos=1``            -> ctc.py:12-14
* ``blank_state``  == ``init_state``                     -> ctc.py:19-27
* ``extend(..., mode="cheap")`` == ``cheap_compute``      -> ctc.py:68-108
* ``extend(..., mode="full")``  == ``full_compute``       -> ctc.py:29-66

All arithmetic is fp32 with ``numpy.logaddexp`` applied in the reference's
order, so results are bit-identical to the reference on the same machine
(checked in tests/test_oracle_vs_reference.py and pinned by tests/golden).
``dtype=np.float64`` gives the error-budget variant.

The recursion, for prefix g (length n, last token l), candidates c_j and
posteriors x[t, v] (log domain), with start = max(1, n):

    sum_prev[t] = logaddexp(r_prev[t,0], r_prev[t,1])
    phi[t, j]   = r_prev[t,1] if (n > 0 and c_j == l) else sum_prev[t]
    r[t,0,j]    = logaddexp(r[t-1,0,j], phi[t-1,j]) + x[t, c_j]        t >= start
    r[t,1,j]    = logaddexp(r[t-1,1,j], r[t-1,0,j]) + x[t, blank]      t >= start
    psi[j]      = logaddexp over t>=start of (phi[t-1,j] + x[t,c_j]), seeded with r[start-1,0,j]
    psi[eos]    = sum_prev[T-1]                                         (cheap mode only)

Rows t < start stay ``logzero`` except r[0,0,:] = x[0, c] for the empty prefix
(and the aliasing quirk noted in ``extend`` when the loop never runs).
"""
import numpy as np

LOGZERO = -100000000.0   # ctc.py:12
BLANK = 0                # ctc.py:13
EOS = 1                  # ctc.py:14


def blank_state(x, dtype=np.float32):
    """State of the empty prefix: r[:,0]=logzero, r[t,1]=sum_{tau<=t} x[tau,blank].

    The running sum is sequential in ``dtype`` exactly like ctc.py:24-26."""
    x = np.asarray(x)
    n_frames = x.shape[0]
    r = np.empty((n_frames, 2), dtype=dtype)
    r[:, 0] = LOGZERO
    acc = dtype(x[0, BLANK])
    r[0, 1] = acc
    for t in range(1, n_frames):
        acc = dtype(acc + dtype(x[t, BLANK]))
        r[t, 1] = acc
    return r


def extend(x, prefix_len, last_tok, r_prev, cands, mode="cheap", dtype=np.float32):
    """Score every one-token extension ``prefix + [c]`` for ``c`` in ``cands``.

    x         [T, V] log posteriors
    prefix_len, last_tok   len(g) and g[-1] of the prefix (last_tok ignored if empty)
    r_prev    [T, 2] state of the prefix (col 0 non-blank-ending, col 1 blank-ending)
    cands     list of token ids (``mode="full"`` ignores it and uses range(V))
    returns   psi [C], r [C, T, 2]   (r is a fresh contiguous array; the reference
              returns the same numbers as a rolled view, ctc.py:108)
    """
    x = np.asarray(x).astype(dtype, copy=False)
    r_prev = np.asarray(r_prev).astype(dtype, copy=False)
    n_frames, vocab = x.shape
    full = (mode == "full")
    if full:
        cands = list(range(vocab))
    else:
        cands = [int(c) for c in cands]
    n_c = len(cands)
    lz = dtype(LOGZERO)

    r = np.full((n_c, n_frames, 2), lz, dtype=dtype)
    start = max(1, prefix_len)
    x_c = x[:, cands]                       # [T, C]  gathered candidate columns
    x_b = x[:, BLANK]                       # [T]
    if prefix_len == 0:
        r[:, 0, 0] = x_c[0]                 # ctc.py:82-83 / :43-44

    both = np.logaddexp(r_prev[:, 0], r_prev[:, 1])          # ctc.py:87
    phi = np.repeat(both[:, None], n_c, axis=1)              # [T, C]
    if full:
        # full_compute masks the non-blank path of column last_char with logzero
        # (ctc.py:53-56); for the empty prefix last_char is 0, i.e. the blank column.
        col = last_tok if prefix_len > 0 else 0
        phi[:, col] = np.logaddexp(np.full(n_frames, lz, dtype=dtype), r_prev[:, 1])
    elif prefix_len > 0 and last_tok in cands:               # ctc.py:90-91
        phi[:, cands.index(last_tok)] = r_prev[:, 1]

    if start - 1 < n_frames:
        psi = r[:, start - 1, 0].copy()                      # ctc.py:85
        nb = r[:, start - 1, 0].copy()
        bl = r[:, start - 1, 1].copy()
    else:
        raise IndexError("prefix longer than the encoder output (ctc.py:85)")
    for t in range(start, n_frames):
        new_nb = np.logaddexp(nb, phi[t - 1]) + x_c[t]       # ctc.py:100
        new_bl = np.logaddexp(bl, nb) + x_b[t]               # ctc.py:102
        psi = np.logaddexp(psi, phi[t - 1] + x_c[t])         # ctc.py:103
        r[:, t, 0] = new_nb
        r[:, t, 1] = new_bl
        nb, bl = new_nb, new_bl

    if not full and EOS in cands:                            # ctc.py:106-107
        j = cands.index(EOS)
        psi[j] = both[-1]
        if start >= n_frames:
            # Reference quirk: psi is a *view* of r[start-1,0,:]
            # (ctc.py:85); when the time loop never runs (len(g) >= T) the eos
            # override above therefore also lands in the returned state.
            r[j, start - 1, 0] = both[-1]
    return psi.astype(dtype, copy=False), r


class PrefixScorerOracle:
    """Object wrapper with the reference's method names, for tests that want to
    read like the reference's call sites (decode.py:96-97,131)."""

    def __init__(self, x, dtype=np.float32):
        x = np.asarray(x)
        if x.ndim == 3:          # [1, T, V] as handed over by decode.py:96
            x = x[0]
        self.x = x.astype(dtype, copy=False)
        self.dtype = dtype
        self.logzero, self.blank, self.eos = LOGZERO, BLANK, EOS
        self.odim = self.x.shape[-1]
        self.input_length = self.x.shape[0]

    def init_state(self):
        return blank_state(self.x, self.dtype)

    def cheap_compute(self, g, r_prev, candidates):
        last = g[-1] if len(g) else 0
        return extend(self.x, len(g), last, r_prev, candidates, "cheap", self.dtype)

    def full_compute(self, g, r_prev):
        last = g[-1] if len(g) else 0
        return extend(self.x, len(g), last, r_prev, None, "full", self.dtype)


def cand_frames(enc_frames, beam, n_cand, n_steps):
    """Unit count of SURVEY.md §8d for one utterance: sum_s H_s * C * T with
    H_0 = 1 and H_s = beam afterwards."""
    if n_steps <= 0:
        return 0
    return (1 + (n_steps - 1) * beam) * n_cand * enc_frames
