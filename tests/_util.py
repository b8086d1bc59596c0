"""Shared helpers for the parity tests (seeded inputs, tolerances, oracle chains)."""
import numpy as np

LOGZERO = -1e8


def log_softmax_np(a):
    m = a.max(-1, keepdims=True)
    return ((a - m) - np.log(np.exp(a - m).sum(-1, keepdims=True))).astype(np.float32)


def posteriors(rng, n_utts, t_max, vocab, scale=2.0):
    """[U, T, V] fp32 log-posteriors shaped like the CTC head's output (ReLU then log-softmax)."""
    lg = rng.standard_normal((n_utts, t_max, vocab)).astype(np.float32) * scale
    return log_softmax_np(np.maximum(lg, 0.0))


def ulp32(v):
    return np.spacing(np.abs(v).astype(np.float32))


def prefix_tolerance(ref, ulps=2.0):
    """north_star: prefix log-probs within 1e-4 absolute in fp32.  One fp32 ulp exceeds 1e-4
    once |value| >= 1024, so beyond that the bound is 2 ulp of the value
    (the opt-in MUFU fast-math variant is held to 4 ulp; it is not the default)."""
    return np.maximum(1e-4, ulps * ulp32(ref))


def assert_prefix_close(got, ref, what, ulps=2.0):
    got, ref = np.asarray(got, np.float32), np.asarray(ref, np.float32)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    live = ref > -1e7                      # log-zero entries drift by multiples of ulp(1e8)=8
    dead_ok = np.all(got[~live] < -1e7)
    err = np.abs(got[live] - ref[live]) if live.any() else np.zeros(1)
    tol = prefix_tolerance(ref[live], ulps) if live.any() else np.ones(1)
    worst = float((err / tol).max()) if live.any() else 0.0
    assert dead_ok, what + ": a log-zero entry of the oracle is finite on the device"
    assert worst <= 1.0, "%s: max |err| %.3g (%.2f x tolerance)" % (what, float(err.max()), worst)
    return float(err.max())
