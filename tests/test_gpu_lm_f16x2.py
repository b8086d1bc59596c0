"""The 2-piece fp16 operand format of the RNNLM step (csrc/lstm_step.cu, kPieces == 2; stepper.SplitLinearF16):
three partial tensor-core products instead of the bf16 format's six.  Same bars as the bf16 format's tests in
test_gpu_kernels.py / test_gpu_decode.py: pieces bit-equal to their torch restatement, cell within 1e-6 of
float64, GEMM error within 2x the library's fp32 GEMM + 1e-6 (measured on the B200: 1.28e-5 vs 1.02e-5 at the
RNNLM's K = 2048, where accumulation dominates; 1.3e-6 vs 0.3e-6 at K = 300, where the operands' 22 bits show),
LSTM stack within 2e-6 of nn.LSTM in float64, and the decode's 1-best identical to the oracle's.
The format is opt-in (BeamDecoder.lm_split = "fp16x2"): at the full workload it changes 1.4 % of the 1-best
sequences against the bf16 format (profiles/r01_compare_lm_fp16x2_vs_bf16x3.json; the batch-shape noise floor is
0.08 %, profiles/r01_compare_control_half_batches.json).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

SCALE = 2.0 ** 14


def _ops():
    from e2e_asr_pytorch_b200 import ops, _lib
    return ops, _lib


@pytest.mark.parametrize("n,w,k,off,gather", [(37, 1024, 2048, 1024, True), (5, 940, 1240, 0, False), (2, 7, 9, 1, True)])
def test_split_rows_f16x2_is_the_two_piece_split(cuda, n, w, k, off, gather):
    ops, _ = _ops()
    from e2e_asr_pytorch_b200.stepper import _split2_f16
    g = torch.Generator().manual_seed(n + w)
    src = torch.tanh(torch.randn(n + 4, w, generator=g) * 2)
    src[0, 0], src[1, 0], src[2, 0] = 1.0, -1.0, 1e-7                   # range ends and a value whose residual is subnormal
    src = src.to(cuda)
    idx = torch.randint(0, n + 4, (n,), generator=g).to(cuda) if gather else None
    dst = torch.full((n + 1, 2 * k), 7.0, dtype=torch.float16, device=cuda)
    ops.lstm_split_rows(src, idx, n, dst, k, off, scale=SCALE)
    rows = src.index_select(0, idx) if gather else src[:n]
    want = _split2_f16(rows, SCALE)
    for p in range(2):
        assert torch.equal(dst[:n, p * k + off:p * k + off + w], want[p])
    mask = torch.ones(2 * k, dtype=torch.bool)
    for p in range(2):
        mask[p * k + off:p * k + off + w] = False
    assert (dst[:n][:, mask.to(cuda)] == 7.0).all() and (dst[n] == 7.0).all()
    back = (want[0].double() + want[1].double()) / SCALE
    assert (back - rows.double()).abs().max().item() <= 2.0 ** -22
    with pytest.raises(Exception):
        ops.lstm_split_rows(src, idx, n, dst, k, off, scale=3.0)         # not a power of two
    with pytest.raises(TypeError):
        ops.lstm_split_rows(src, idx, n, dst, k, off)                    # fp16 operand without a scale


@pytest.mark.parametrize("n,d,with_table,with_next", [(33, 1024, True, True), (9, 300, False, False), (4, 12, True, True)])
def test_lstm_cell_f16x2_matches_torch(cuda, n, d, with_table, with_next):
    ops, _ = _ops()
    from e2e_asr_pytorch_b200.stepper import _split2_f16
    g = torch.Generator().manual_seed(d)
    gate_scale = 2.0 ** -27
    gates = torch.randn(n, 4 * d, generator=g) * 2
    bias = torch.randn(4 * d, generator=g)
    table = torch.randn(5, 4 * d, generator=g) if with_table else None
    tok = torch.randint(0, 5, (n,), generator=g) if with_table else None
    c_prev = torch.randn(n + 3, d, generator=g)
    idx = torch.randint(0, n + 3, (n,), generator=g)
    z = gates.double() + bias.double() + (table.double()[tok] if with_table else 0)
    i, f, gg, o = z.chunk(4, dim=-1)
    c_want = torch.sigmoid(f) * c_prev.double()[idx] + torch.sigmoid(i) * torch.tanh(gg)
    h_want = torch.sigmoid(o) * torch.tanh(c_want)
    dev = lambda t: None if t is None else t.to(cuda)
    c_new, h_new = torch.empty(n, d, device=cuda), torch.empty(n, d, device=cuda)
    k_next = 2 * d
    a_next = torch.zeros(n, 2 * k_next, dtype=torch.float16, device=cuda) if with_next else None
    ops.lstm_cell(dev(gates / gate_scale), dev(bias), dev(c_prev), dev(idx), n, c_new, h_new, table=dev(table), tok=dev(tok),
                  a_next=a_next, k_next=k_next if with_next else 0, off_next=0, gate_scale=gate_scale, next_scale=SCALE)
    assert (c_new.cpu().double() - c_want).abs().max().item() < 1e-6
    assert (h_new.cpu().double() - h_want).abs().max().item() < 1e-6
    if with_next:
        want = _split2_f16(h_new, SCALE)
        for p in range(2):
            assert torch.equal(a_next[:, p * k_next:p * k_next + d], want[p])


def test_f16x2_gemm_is_fp32_accurate(cuda):
    """x W^T from the three fp16 partial products vs float64: its error must not exceed that of the library's own fp32
    GEMM on the same data (the bar SplitLinear is held to in test_gpu_decode.py), for hidden-state-like inputs."""
    from e2e_asr_pytorch_b200.stepper import SplitLinearF16
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    g = torch.Generator().manual_seed(0)
    with _Fp32Math():
        for n, k, m in [(2048, 2048, 4096), (1536, 1024, 4096), (8, 300, 32)]:
            x = torch.tanh(torch.randn(n, k, generator=g) * 3).to(cuda)
            w = (torch.randn(m, k, generator=g) / k ** 0.5).to(cuda)
            want = x.double() @ w.double().t()
            e_split = (SplitLinearF16(w)(x).double() - want).abs().max().item()
            e_fp32 = (torch.nn.functional.linear(x, w).double() - want).abs().max().item()
            print("fp16x2 gemm %dx%dx%d: max err %.3g (cuBLAS fp32 %.3g)" % (n, k, m, e_split, e_fp32))
            # short contractions show the operands' 22 bits (measured 1.3e-6 at K = 300) rather than accumulation error
            assert e_split <= max(2.0 * e_fp32 + 1e-6, 3e-6)


def test_fused_lstm_stack_f16x2_matches_nn_lstm(cuda):
    _ops()
    from e2e_asr_pytorch_b200.stepper import _FusedLstm
    from e2e_asr_pytorch_b200.decode import _Fp32Math
    torch.manual_seed(3)
    for in_dim, d, layers in [(16, 16, 3), (32, 64, 4)]:
        rnn = torch.nn.LSTM(in_dim, d, num_layers=layers, batch_first=True)
        emb = torch.randn(7, in_dim)
        ref = torch.nn.LSTM(in_dim, d, num_layers=layers, batch_first=True).double()
        ref.load_state_dict({k: v.double() for k, v in rnn.state_dict().items()})
        n = 10
        with torch.no_grad(), _Fp32Math():
            fused = _FusedLstm(rnn.to(cuda), emb.to(cuda), split="fp16x2")
            fused.start(n, cuda)
            h = torch.zeros(layers, n, d, dtype=torch.float64)
            c = torch.zeros(layers, n, d, dtype=torch.float64)
            g = torch.Generator().manual_seed(5)
            for step in range(4):
                tok = torch.randint(0, 7, (n,), generator=g)
                top = fused.step(n, tok=tok.to(cuda))
                out, (h, c) = ref(emb[tok].double()[:, None, :], (h, c))
                assert (top.cpu().double() - out[:, 0]).abs().max().item() < 2e-6
                perm = torch.randint(0, n, (n,), generator=g)
                fused.reorder(perm.to(cuda))
                h, c = h[:, perm], c[:, perm]
                assert (fused.hidden(n).cpu().double() - torch.cat(list(h), dim=1)).abs().max().item() < 2e-6
    with pytest.raises(NotImplementedError):
        _FusedLstm(rnn, None, split="fp16x2")                              # untabled layer 0: inputs of unknown range


def test_decode_with_f16x2_lm_matches_oracle(cuda):
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    from tests.test_gpu_decode import _models, _oracle_nbest, _compare
    asr, lm, lm_path, lm_cfg = _models()
    lens = [64, 120, 92, 200, 76, 148]
    feat, fl = synth.padded_batch(list(range(len(lens))), lens)
    dec = BeamDecoder(asr, None, 8, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=0.5, ctc_weight=0.5).to(cuda)
    dec.lm_split = "fp16x2"
    out = dec.decode_batch(feat.to(cuda), fl.to(cuda))
    from e2e_asr_pytorch_b200.stepper import _FusedLstm
    assert isinstance(dec._stepper[2].lm_rnn, _FusedLstm) and dec._stepper[2].lm_rnn.f16
    same = ties = 0
    for k, n in enumerate(lens):
        ora = _oracle_nbest(asr, lm, feat[k], n, 8, 0.5, 0.5)
        assert len(out[k]) == len(ora)
        s, t = _compare(out[k], ora, "fp16x2 lm utt %d" % k)
        same, ties = same + s, ties + t
    print("fp16x2 LM: identical 1-best %d/%d, ties %d" % (same, len(lens), ties))
    assert same >= len(lens) - 1
