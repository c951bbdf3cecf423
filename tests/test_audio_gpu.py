"""CUDA STFT / mel front end vs the oracle (torch.stft on CPU).  Floating point: tolerances stated per test."""
import numpy as np
import pytest
import torch

import vits_oracle as O

pytestmark = pytest.mark.gpu


def _wav(b, t, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 1, t, generator=g) * 2 - 1
    tt = torch.arange(t) / 22050.0
    return (0.5 * x + 0.4 * torch.sin(2 * np.pi * 440 * tt)[None, None]).clamp(-1, 1)


@pytest.mark.parametrize("t", [8192, 41885, 1500])
def test_wav_to_spec_and_mel_vs_oracle(t):
    from tts_b200 import audio
    wav = _wav(2, t, t)
    spec = audio.wav_to_spec(wav.cuda(), 1024, 256, 1024)
    want = O.wav_to_spec(wav, 1024, 256, 1024)
    assert spec.shape == want.shape
    # magnitudes reach ~200; 1e-5 relative to the peak is the reference's own TorchSTFT-vs-librosa bound
    assert (spec.cpu() - want).abs().max() <= 1e-5 * max(1.0, want.max().item()) + 2e-4
    mel = audio.spec_to_mel(spec, 1024, 80, 22050, 0, None)
    want_mel = O.spec_to_mel(want, 1024, 80, 22050, 0, None)
    assert (mel.cpu() - want_mel).abs().max() <= 2e-3   # log of small band energies amplifies the 1e-5 error
    # the relation the reference pins exactly (tests/tts_tests/test_vits.py:56)
    assert torch.equal(audio.wav_to_mel(wav.cuda(), 1024, 80, 22050, 256, 1024, 0, None), mel)


def test_short_window_and_other_sizes():
    from tts_b200 import audio
    wav = _wav(3, 5000, 9)
    got = audio.wav_to_spec(wav.cuda(), 512, 128, 400)
    want = O.wav_to_spec(wav, 512, 128, 400)
    assert got.shape == want.shape
    assert (got.cpu() - want).abs().max() <= 1e-5 * want.max().item() + 2e-4


def test_torch_stft_class_vs_oracle():
    from tts_b200.audio import TorchSTFT, mel_filterbank
    wav = _wav(2, 7000, 4).squeeze(1)
    stft = TorchSTFT(n_fft=1024, hop_length=256, win_length=1024)
    got = stft(wav.cuda())
    want = O.torch_stft_call(wav, 1024, 256, 1024)
    assert got.shape == want.shape
    assert (got.cpu() - want).abs().max() <= 1e-5 * want.max().item() + 2e-4
    stft2 = TorchSTFT(1024, 256, 1024, pad_wav=True, sample_rate=22050, n_mels=80, use_mel=True, do_amp_to_db=True)
    got2 = stft2(wav.cuda())
    basis = torch.from_numpy(O.slaney_mel_basis(22050, 1024, 80, 0, None))
    want2 = O.torch_stft_call(wav, 1024, 256, 1024, pad_wav=True, use_mel=True, mel_basis=basis, do_amp_to_db=True)
    assert got2.shape == want2.shape
    assert (got2.cpu() - want2).abs().max() <= 2e-3
    assert np.abs(mel_filterbank(22050, 1024, 80, 0, None) - basis.numpy()).max() < 1e-7
