"""Drop-in ``CTCPrefixScore`` backed by the sm_100a prefix-score kernel.

Mirrors the interface of ``/root/reference/src/ctc.py`` (class ``CTCPrefixScore``:
``__init__(x)`` :11-17, ``init_state()`` :19-27, ``full_compute(g, r_prev)`` :29-66,
``cheap_compute(g, r_prev, candidates)`` :68-108; attributes ``logzero``, ``blank``,
``eos``, ``x``, ``odim``, ``input_length``) so call sites such as
``src/decode.py:96-97,131`` work unchanged.  Every call is one launch of
``e2e_ctc_prefix_score`` with U=1, B=1; there is no CPU path.

By default results come back as numpy arrays with the reference's shapes
(``psi [C]`` fp32, ``r [C,T,2]`` fp32).  ``device_state=True`` keeps them as CUDA
tensors instead (same shapes) to avoid the device->host copy the reference's
numpy contract forces on every call.
"""
import numpy as np
import torch

from . import _lib as L
from . import ops


class CTCPrefixScore:
    def __init__(self, x, device_state=False, fast_math=False):
        if not torch.is_tensor(x):
            raise TypeError("CTCPrefixScore expects the [1,T,V] log-posterior tensor (src/decode.py:94-96)")
        if not x.is_cuda:
            raise L.E2EError("CTCPrefixScore has no CPU path: move the posteriors to a CUDA device")
        self.logzero = -100000000.0
        self.blank = 0
        self.eos = 1
        self._xt = x[0].detach().to(torch.float32)            # [T, V] (batch item 0, like src/ctc.py:15)
        self.odim = x.shape[-1]
        self.input_length = self._xt.shape[0]
        self._device_state = device_state
        self._flags = L.PREFIX_FAST_MATH if fast_math else 0
        t, v = self._xt.shape
        vp = ops.padded_vocab(v)
        xp = torch.full((t, 1, vp), self.logzero, dtype=torch.float32, device=x.device)
        xp[:, 0, :v] = self._xt
        self._x_dev = xp                                       # [T, 1, Vp] frame-major, U = 1
        self._enc_len = torch.tensor([t], dtype=torch.int32, device=x.device)
        self._one = torch.ones(1, dtype=torch.int32, device=x.device)
        self._zero = torch.zeros(1, dtype=torch.int32, device=x.device)
        self._x_np = None

    @property
    def x(self):
        """numpy view of the posteriors, as the reference exposes it (src/ctc.py:15)."""
        if self._x_np is None:
            self._x_np = self._xt.cpu().numpy()
        return self._x_np

    def _wrap(self, t):
        return t if self._device_state else t.cpu().numpy()

    def init_state(self):
        r0 = ops.ctc_init_state(self._x_dev, self._enc_len)    # [1, T, 1, 2]
        return self._wrap(r0.view(self.input_length, 2))

    def _run(self, g, r_prev, candidates, full):
        dev = self._x_dev.device
        t = self.input_length
        if len(g) > t:
            # same failure the reference hits at ``psi = r[start-1, 0, :]`` (src/ctc.py:85)
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (len(g) - 1, t))
        rp = torch.as_tensor(r_prev, dtype=torch.float32).to(dev).reshape(1, t, 1, 2).contiguous()
        if full:
            c = self.odim
            cand = None
        else:
            candidates = [int(v) for v in candidates]
            c = len(candidates)
            cand = torch.tensor(candidates, dtype=torch.int32, device=dev).view(1, c)
        last = torch.tensor([int(g[-1]) if len(g) else 0], dtype=torch.int32, device=dev)
        plen = torch.tensor([len(g)], dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        flags = self._flags | (L.PREFIX_FULL if full else 0)
        psi, r = ops.ctc_prefix_score(self._x_dev, self.odim, self._enc_len, rp, self._zero, last, plen,
                                      self._one, cand, 1, c, flags, status=status)
        # r_out is [1, T, C, 2]; the reference hands back np.rollaxis(r[T,2,C], 2) = [C, T, 2]
        return self._wrap(psi.view(c)), self._wrap(r.view(t, c, 2).permute(1, 0, 2))

    def cheap_compute(self, g, r_prev, candidates):
        return self._run(g, r_prev, candidates, full=False)

    def full_compute(self, g, r_prev):
        return self._run(g, r_prev, None, full=True)
