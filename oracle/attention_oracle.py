"""CPU restatement of one location-aware attention step (TEST INFRASTRUCTURE).

Follows ``/root/reference/src/module.py``: ``LocationAwareAttention.forward`` (:1152-1173, minus
the convolution, whose output is an input here) and ``BaseAttention._attend`` (:1109-1117, minus the
context product), with ``compute_mask`` (:1100-1107).  Batched over hypotheses n = u*B + b; checked
against the product's own reference-shaped module in tests/test_oracle_golden.py and used as the
checker of ``e2e_attention_loc_full`` in tests/test_gpu_kernels.py.
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


def loc_attention_full(key, value, query, prev_att, enc_len, w_conv, w_proj, w_energy, b_energy, temperature, beam):
    """The whole step, module.py:1152-1173 + :1109-1117: conv (Conv1d 1->K, zero padded, no bias) ->
    :func:`loc_attention_step` -> context.  value [U,T,E], prev_att [N,T], w_conv [K,W] -> (attn [N,T], ctx [N,E])."""
    prev_att = torch.as_tensor(prev_att, dtype=torch.float32)
    w_conv = torch.as_tensor(w_conv, dtype=torch.float32)
    value = torch.as_tensor(value, dtype=torch.float32)
    pad = w_conv.shape[1] // 2
    feat = torch.nn.functional.conv1d(prev_att[:, None, :], w_conv[:, None, :], padding=pad)     # module.py:1163
    attn = loc_attention_step(key, query, feat, enc_len, w_proj, w_energy, b_energy, temperature, beam)
    u_of = torch.arange(attn.shape[0]) // beam
    ctx = torch.bmm(attn[:, None, :], value[u_of]).squeeze(1)                                      # module.py:1114
    return attn, ctx
