"""Batched (utterance x beam) model step in plain PyTorch.

north_star keeps the encoder, the location-aware attention decoder step and the
RNNLM step as cuBLAS/cuDNN-backed PyTorch.  The reference runs them one
hypothesis at a time on batch-1 tensors and bounces every state through the
CPU (src/decode.py:105-123,144-151,265-277); here the same arithmetic runs once
per decode step over all N = U*B hypotheses with device-resident states:

* attention  (src/asr.py:333-364, src/module.py:1152-1173 / :1120-1132)
* speller    (src/asr.py:259-266)
* RNNLM      (src/lm.py:27-38)

The stepper reads the weights straight out of the ``asr`` / ``lm`` modules it is
given (parameter names of the reference), so it works with the reference's own
objects as well as with ``model.py``.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _rnn_weights(rnn):
    ws = []
    for l in range(rnn.num_layers):
        ws.append(tuple(getattr(rnn, "{}_l{}".format(k, l)) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")))
    return ws


def _lstm_cell(x, h, c, w):
    w_ih, w_hh, b_ih, b_hh = w
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, dim=-1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def _gru_cell(x, h, w):
    w_ih, w_hh, b_ih, b_hh = w
    gi, gh = F.linear(x, w_ih, b_ih), F.linear(h, w_hh, b_hh)
    ir, iz, inn = gi.chunk(3, dim=-1)
    hr, hz, hn = gh.chunk(3, dim=-1)
    r, z = torch.sigmoid(ir + hr), torch.sigmoid(iz + hz)
    n = torch.tanh(inn + r * hn)
    return (1 - z) * n + z * h


class _Rnn:
    """n-layer LSTM/GRU advanced one token at a time for a batch of rows."""

    def __init__(self, rnn):
        self.is_lstm = isinstance(rnn, torch.nn.LSTM)
        if not self.is_lstm and not isinstance(rnn, torch.nn.GRU):
            raise NotImplementedError("only LSTM / GRU recurrent layers are supported")
        self.w = _rnn_weights(rnn)
        self.layers, self.dim = rnn.num_layers, rnn.hidden_size

    def zeros(self, n, device):
        h = torch.zeros(self.layers, n, self.dim, device=device)
        return (h, torch.zeros_like(h)) if self.is_lstm else (h, None)

    def step(self, x, state):
        h, c = state
        hs, cs = [], []
        for l in range(self.layers):
            if self.is_lstm:
                x, c2 = _lstm_cell(x, h[l], c[l], self.w[l])
                cs.append(c2)
            else:
                x = _gru_cell(x, h[l], self.w[l])
            hs.append(x)
        return x, (torch.stack(hs), torch.stack(cs) if self.is_lstm else None)

    @staticmethod
    def gather(state, idx):
        h, c = state
        return (h.index_select(1, idx), c.index_select(1, idx) if c is not None else None)


class BatchedStepper:
    def __init__(self, asr, lm=None):
        att = asr.attention
        if att.num_head != 1:
            raise NotImplementedError("multi-head attention is not supported by the batched beam search")
        if att.mode not in ("loc", "dot"):
            raise NotImplementedError("attention mode " + str(att.mode))
        self.asr, self.lm = asr, lm
        self.mode, self.temperature = att.mode, att.att_layer.temperature
        self.dec = _Rnn(asr.decoder.layers)
        self.lm_rnn = _Rnn(lm.rnn) if lm is not None else None

    # -- once per batch ---------------------------------------------------------------------
    def encode(self, feats, lens):
        """feats [U,Lmax,D] zero padded, lens [U] -> enc [U,Tmax,E], enc_len [U] (long, same device)."""
        enc_mod = self.asr.encoder
        if feats.shape[0] == 1:
            enc, enc_len = enc_mod(feats, lens)
        elif hasattr(enc_mod, "forward_ragged"):
            enc, enc_len = enc_mod.forward_ragged(feats, lens)
        else:   # opaque encoder (e.g. the reference's own module): one exact batch-1 call per utterance
            outs, ls = [], []
            for i in range(feats.shape[0]):
                n = int(lens[i])
                e, l = enc_mod(feats[i:i + 1, :n], lens[i:i + 1])
                outs.append(e[0])
                ls.append(l.reshape(-1)[0])
            enc = torch.nn.utils.rnn.pad_sequence(outs, batch_first=True)
            enc_len = torch.stack(ls)
        enc_len = enc_len.to(enc.device).long().clamp(max=enc.shape[1])
        t_max = int(enc_len.max())
        return enc[:, :t_max].contiguous(), enc_len

    def start(self, enc, enc_len, beam):
        att = self.asr.attention
        u, t, _ = enc.shape
        self.U, self.B, self.T, self.N = u, beam, t, u * beam
        dev = enc.device
        self.key = torch.tanh(att.proj_k(enc))                                   # [U,T,A]   asr.py:343
        self.value = torch.tanh(att.proj_v(enc)) if att.v_proj else enc          # [U,T,E]
        self.pad = torch.arange(t, device=dev)[None, :] >= enc_len[:, None]      # [U,T]     module.py:1100-1107
        if self.mode == "loc":
            uni = (1.0 / enc_len.to(torch.float32))[:, None].expand(u, t).masked_fill(self.pad, 0.0)   # module.py:1157-1160
            self.prev_att = uni[:, None, :].expand(u, beam, t).reshape(self.N, t).contiguous()
        else:
            self.prev_att = None
        self.dec_state = self.dec.zeros(self.N, dev)
        self.lm_state = self.lm_rnn.zeros(self.N, dev) if self.lm_rnn is not None else None

    # -- once per decode step ---------------------------------------------------------------
    def step(self, prev_tok):
        """prev_tok [N] long -> (att_logits [N,V], lm_logits [N,V] | None); new states are held
        until :meth:`reorder` commits them in the surviving hypotheses' order."""
        asr, att = self.asr, self.asr.attention
        u, b, t = self.U, self.B, self.T
        h = self.dec_state[0]
        query = torch.tanh(att.proj_q(h.transpose(0, 1).reshape(self.N, -1)))     # asr.py:337 / :251-254
        if self.mode == "loc":
            lay = att.att_layer
            loc = lay.loc_conv(self.prev_att[:, None, :]).transpose(1, 2)          # [N,T,K]
            loc = torch.tanh(lay.loc_proj(loc)).view(u, b, t, -1)
            mix = torch.tanh(self.key[:, None] + query.view(u, b, 1, -1) + loc)
            energy = lay.gen_energy(mix).squeeze(-1)                               # [U,B,T]  module.py:1168
        else:
            energy = torch.bmm(query.view(u, b, -1), self.key.transpose(1, 2))     # module.py:1126
        score = (energy / self.temperature).masked_fill(self.pad[:, None, :], -np.inf)
        attn = torch.softmax(score, dim=-1)                                        # [U,B,T]
        context = torch.bmm(attn, self.value).view(self.N, -1)                     # module.py:1114
        dec_in = torch.cat([asr.pre_embed(prev_tok), context], dim=-1)             # decode.py:114-115
        top, self._new_dec = self.dec.step(dec_in, self.dec_state)
        att_logits = asr.decoder.char_trans(top)                                   # asr.py:265
        self._new_att = attn.view(self.N, t) if self.mode == "loc" else None
        lm_logits = None
        if self.lm is not None:
            top, self._new_lm = self.lm_rnn.step(self.lm.emb(prev_tok), self.lm_state)
            lm_logits = F.linear(top, self.lm.emb.weight) if self.lm.emb_tying else self.lm.trans(top)   # lm.py:33-37
        return att_logits.contiguous(), (lm_logits.contiguous() if lm_logits is not None else None)

    def reorder(self, parent_slot):
        """parent_slot [U,B] int32 (slot of each survivor's parent): children inherit the
        post-step states of their parent (decode.py:159-162,250-257)."""
        base = torch.arange(self.U, device=parent_slot.device, dtype=torch.long)[:, None] * self.B
        idx = (base + parent_slot.long()).reshape(-1)
        self.dec_state = _Rnn.gather(self._new_dec, idx)
        if self._new_att is not None:
            self.prev_att = self._new_att.index_select(0, idx)
        if self.lm is not None:
            self.lm_state = _Rnn.gather(self._new_lm, idx)
