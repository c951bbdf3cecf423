"""Drop-ins for the STFT / mel front end on the hot path:
``wav_to_spec`` / ``spec_to_mel`` / ``wav_to_mel`` / ``amp_to_db`` / ``db_to_amp``
(/root/reference/TTS/tts/models/vits.py:78-208) and ``TorchSTFT``
(/root/reference/TTS/utils/audio/torch_transforms.py:6-165), computed by libtts_b200.so.

The mel filterbank is the Slaney-style ``librosa.filters.mel`` (htk=False, norm="slaney") the reference
calls; librosa is not a dependency here, ``mel_filterbank`` below builds the same matrix on the host.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib

_handles = {}


def mel_filterbank(sample_rate, n_fft, n_mels, fmin=0.0, fmax=None):
    """Triangular mel filters on the Slaney scale (linear below 1 kHz, log above), area-normalised."""
    fmax = float(sample_rate) / 2.0 if fmax is None else float(fmax)
    lin_step, knee_hz = 200.0 / 3.0, 1000.0
    knee_mel, log_step = knee_hz / lin_step, math.log(6.4) / 27.0

    def to_mel(hz):
        hz = np.asarray(hz, dtype=np.float64)
        return np.where(hz >= knee_hz, knee_mel + np.log(np.maximum(hz, 1e-10) / knee_hz) / log_step, hz / lin_step)

    def to_hz(mel):
        mel = np.asarray(mel, dtype=np.float64)
        return np.where(mel >= knee_mel, knee_hz * np.exp(log_step * (mel - knee_mel)), mel * lin_step)

    edges = to_hz(np.linspace(to_mel(fmin), to_mel(fmax), n_mels + 2))
    bins = np.fft.rfftfreq(n_fft, d=1.0 / sample_rate)
    width = np.diff(edges)
    rel = edges[:, None] - bins[None, :]
    fb = np.zeros((n_mels, bins.shape[0]), dtype=np.float32)
    for m in range(n_mels):
        rising = -rel[m] / width[m]
        falling = rel[m + 2] / width[m + 1]
        fb[m] = np.maximum(0.0, np.minimum(rising, falling))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return fb


def _window(win_length, n_fft, name="hann_window"):
    w = getattr(torch, name)(win_length).to(torch.float32)
    if win_length < n_fft:  # torch.stft centres a short window inside the frame
        left = (n_fft - win_length) // 2
        w = torch.nn.functional.pad(w, (left, n_fft - win_length - left))
    return w.contiguous()


def _handle(device, n_fft, hop, win_length, window_name, mel_key=None, mel_basis=None):
    key = (str(device), n_fft, hop, win_length, window_name, mel_key)
    h = _handles.get(key)
    if h is None:
        w = _window(win_length, n_fft, window_name)
        mb = None if mel_basis is None else torch.as_tensor(mel_basis, dtype=torch.float32).contiguous()
        out = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = _lib.lib().b200tts_stft_create(n_fft, hop, _lib.ptr(w), _lib.ptr(mb),
                                                0 if mb is None else mb.shape[0], ctypes.byref(out))
        _lib.check(rc, "stft_create")
        _handles[key] = h = out
    return h


def _magnitude(h, y, n_fft, hop, pad1, pad2, mode, power=1.0):
    _lib.require_cuda(y, "y")
    y = y.to(torch.float32).contiguous()
    b, t = y.shape
    n_frames = 1 + (t + 2 * pad1 + 2 * pad2 - n_fft) // hop
    spec = torch.empty((b, n_fft // 2 + 1, max(n_frames, 0)), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        rc = _lib.lib().b200tts_stft_magnitude(h, _lib.ptr(y), b, t, pad1, pad2, mode, ctypes.c_float(power),
                                               _lib.ptr(spec), n_frames, _lib.stream_ptr(y.device))
    _lib.check(rc, "stft_magnitude")
    return spec


def _project(h, spec, n_mels, log_clamp):
    spec = spec.contiguous()
    b, _, n = spec.shape
    mel = torch.empty((b, n_mels, n), dtype=torch.float32, device=spec.device)
    with torch.cuda.device(spec.device):
        rc = _lib.lib().b200tts_stft_mel_project(h, _lib.ptr(spec), b, n, ctypes.c_float(log_clamp), _lib.ptr(mel),
                                                 _lib.stream_ptr(spec.device))
    _lib.check(rc, "stft_mel_project")
    return mel


def amp_to_db(magnitudes, C=1, clip_val=1e-5):
    return torch.log(torch.clamp(magnitudes, min=clip_val) * C)


def db_to_amp(magnitudes, C=1):
    return torch.exp(magnitudes) / C


def wav_to_spec(y, n_fft, hop_length, win_length, center=False):
    """y [B,1,T] -> linear magnitude spectrogram [B, n_fft/2+1, frames]   (vits.py:96-138)."""
    if center:
        raise NotImplementedError("tts_b200.wav_to_spec: the reference only calls this with center=False")
    y = y.squeeze(1)
    h = _handle(y.device, n_fft, hop_length, win_length, "hann_window")
    return _magnitude(h, y, n_fft, hop_length, int((n_fft - hop_length) / 2), 0, mode=0)


def spec_to_mel(spec, n_fft, num_mels, sample_rate, fmin, fmax):
    """[B, n_fft/2+1, T] -> log-mel [B, num_mels, T]   (vits.py:141-157)."""
    _lib.require_cuda(spec, "spec")
    basis = mel_filterbank(sample_rate, n_fft, num_mels, fmin, fmax)
    h = _handle(spec.device, n_fft, 1, n_fft, "hann_window", mel_key=(sample_rate, num_mels, fmin, fmax), mel_basis=basis)
    return _project(h, spec.to(torch.float32), num_mels, 1e-5)


def wav_to_mel(y, n_fft, num_mels, sample_rate, hop_length, win_length, fmin, fmax, center=False):
    """y [B,1,T] -> log-mel [B, num_mels, frames]   (vits.py:160-208)."""
    return spec_to_mel(wav_to_spec(y, n_fft, hop_length, win_length, center), n_fft, num_mels, sample_rate, fmin, fmax)


class TorchSTFT(torch.nn.Module):
    """Same constructor and call contract as TTS.utils.audio.torch_transforms.TorchSTFT."""

    def __init__(self, n_fft, hop_length, win_length, pad_wav=False, window="hann_window", sample_rate=None,
                 mel_fmin=0, mel_fmax=None, n_mels=80, use_mel=False, do_amp_to_db=False, spec_gain=1.0, power=None,
                 use_htk=False, mel_norm="slaney", normalized=False):
        super().__init__()
        if use_htk or mel_norm != "slaney" or normalized:
            raise NotImplementedError("tts_b200.TorchSTFT: only the Slaney mel / un-normalised STFT is built")
        self.n_fft, self.hop_length, self.win_length, self.pad_wav = n_fft, hop_length, win_length, pad_wav
        self.sample_rate, self.mel_fmin, self.mel_fmax, self.n_mels = sample_rate, mel_fmin, mel_fmax, n_mels
        self.use_mel, self.do_amp_to_db, self.spec_gain, self.power = use_mel, do_amp_to_db, spec_gain, power
        self.window_name = window
        self.window = torch.nn.Parameter(getattr(torch, window)(win_length), requires_grad=False)
        self.mel_basis = None
        if use_mel:
            self.mel_basis = torch.from_numpy(mel_filterbank(sample_rate, n_fft, n_mels, mel_fmin, mel_fmax)).float()

    def __call__(self, x):
        """x [B,T] or [B,1,T] -> [B, n_fft/2+1 or n_mels, frames]   (torch_transforms.py:104-145)."""
        if x.ndim == 3:
            x = x.squeeze(1)
        key = (self.sample_rate, self.n_mels, self.mel_fmin, self.mel_fmax) if self.use_mel else None
        h = _handle(x.device, self.n_fft, self.hop_length, self.win_length, self.window_name, mel_key=key,
                    mel_basis=None if self.mel_basis is None else self.mel_basis.numpy())
        pad1 = int((self.n_fft - self.hop_length) / 2) if self.pad_wav else 0
        s = _magnitude(h, x, self.n_fft, self.hop_length, pad1, self.n_fft // 2, mode=1,
                       power=1.0 if self.power is None else float(self.power))
        if self.use_mel:
            s = _project(h, s, self.n_mels, 0.0)
        if self.do_amp_to_db:
            s = torch.log(torch.clamp(s, min=1e-5) * self.spec_gain)
        return s
