"""CPU restatement of one location-aware attention step (TEST INFRASTRUCTURE).

Follows ``/root/reference/src/module.py``: ``LocationAwareAttention.forward`` (:1152-1173, minus
the convolution, whose output is an input here) and ``BaseAttention._attend`` (:1109-1117, minus the
context product), with ``compute_mask`` (:1100-1107).  Batched over hypotheses n = u*B + b; checked
against the product's own reference-shaped module in tests/test_oracle_golden.py and used as the
checker of ``e2e_attention_loc_step`` in tests/test_gpu_kernels.py.
"""
import numpy as np
import torch


def loc_attention_step(key, query, loc_feat, enc_len, w_proj, w_energy, b_energy, temperature, beam):
    """key [U,T,A], query [N,A], loc_feat [N,K,T], enc_len [U], w_proj [A,K], w_energy [A] -> attn [N,T]."""
    key, query, loc_feat = (torch.as_tensor(a, dtype=torch.float32) for a in (key, query, loc_feat))
    w_proj, w_energy = torch.as_tensor(w_proj, dtype=torch.float32), torch.as_tensor(w_energy, dtype=torch.float32)
    n, _, t = loc_feat.shape
    u_of = torch.arange(n) // beam
    loc = torch.tanh(loc_feat.transpose(1, 2) @ w_proj.t())                       # module.py:1163  [N,T,A]
    mix = torch.tanh(key[u_of] + query[:, None, :] + loc)                         # module.py:1168
    energy = mix @ w_energy + b_energy                                            # [N,T]
    pad = torch.arange(t)[None, :] >= torch.as_tensor(enc_len).long()[u_of][:, None]
    score = (energy / temperature).masked_fill(pad, -np.inf)                      # module.py:1110-1111
    return torch.softmax(score, dim=-1)                                           # module.py:1112
