"""BeamDecoder.decode_batch_from_host: pinned host features in, only the valid frames of every utterance copied —
the N-best must equal decode_batch on the same features already resident, whatever the host tensor holds in its
padding."""

import pytest
import torch

pytestmark = pytest.mark.gpu      # confirmed on a B200 (profiles/r02_a_pytest_gpu.txt)


def test_decode_batch_from_host_equals_the_resident_path(cuda):
    from e2e_asr_pytorch_b200 import BeamDecoder, synth
    from tests.test_gpu_decode import _models
    asr, lm, lm_path, lm_cfg = _models()
    lens = [200, 148, 120, 92, 76, 64]
    feat, fl = synth.padded_batch(list(range(len(lens))), lens)
    dec = BeamDecoder(asr, None, 4, 0.01, 0.2, lm_path=lm_path, lm_config=lm_cfg, lm_weight=0.3, ctc_weight=0.5).to(cuda)
    want = dec.decode_batch(feat.to(cuda), fl.to(cuda), return_arrays=True)
    dirty = feat.clone()
    for u, n in enumerate(lens):
        dirty[u, n:] = 1e30                                     # padding that must never reach the device
    got = dec.decode_batch_from_host(dirty.pin_memory(), fl, cuda, return_arrays=True)
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    assert dec.last_h2d_bytes == sum(lens) * feat.shape[2] * 4 + len(lens) * 8
    with pytest.raises(ValueError):
        dec.decode_batch_from_host(feat.clone(), fl, cuda)      # pageable memory
    same = dec.decode_batch_from_host(feat.to(cuda), fl.to(cuda), cuda, return_arrays=True)     # device tensors pass through
    for a, b in zip(want, same):
        assert torch.equal(a, b)
