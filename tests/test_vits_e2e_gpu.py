"""End-to-end Vits.inference on CUDA vs the oracle's restated glue (BASELINE config 2 / 5 shapes, shortened)."""
import pytest
import torch

import vits_oracle as O

pytestmark = pytest.mark.gpu


def _perturb(m, seed):
    gen = torch.Generator().manual_seed(seed)
    for _, p in m.named_parameters():
        if float(p.abs().sum()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=gen) * 0.05)


def _args_dict(a):
    from dataclasses import asdict
    return asdict(a)


def _run(cfg, b, t, seed, lengths=None, speaker_ids=None, length_scale=1.0):
    from tts_b200.vits import Vits
    torch.manual_seed(seed)
    m = Vits(cfg).eval()
    _perturb(m, seed)
    m.length_scale = length_scale
    a = _args_dict(cfg.model_args)
    a["length_scale"] = length_scale
    tok = torch.randint(0, cfg.model_args.num_chars, (b, t))
    lens = torch.full((b,), t) if lengths is None else lengths
    sdp_noise = torch.randn(b, 2, t)
    noise_store = {}

    def prior_noise(shape):
        gen = torch.Generator().manual_seed(seed + 99)
        noise_store["n"] = torch.randn(shape, generator=gen)
        return noise_store["n"]

    want = O.vits_inference(m.state_dict(), tok, lens, sdp_noise, prior_noise, args=a, speaker_ids=speaker_ids)
    m.cuda()
    aux = {"x_lengths": lens.cuda(), "speaker_ids": None if speaker_ids is None else speaker_ids.cuda()}
    got = m.inference(tok.cuda(), aux, sdp_noise=sdp_noise, prior_noise=lambda s: noise_store["n"].cuda())
    return got, want


def _check(got, want):
    # bit-exact integer-valued outputs (north_star: durations / path indices)
    assert torch.equal(got["durations"].cpu(), want["durations"]), "durations differ"
    assert torch.equal(got["y_lengths"].cpu(), want["y_lengths"])
    assert torch.equal(got["alignments"].cpu(), want["alignments"]), "alignment path differs"
    assert torch.equal(got["y_mask"].cpu(), want["y_mask"])
    for k in ("m_p", "logs_p", "z_p", "z"):
        err = (got[k].cpu() - want[k]).abs().max().item()
        assert err < 2e-4, (k, err)
    wav_err = got["model_outputs"].cpu() - want["model_outputs"]
    rms = wav_err.pow(2).mean().sqrt().item()
    assert got["model_outputs"].shape == want["model_outputs"].shape
    assert rms <= 1e-4, f"waveform RMS error {rms} (north_star bound 1e-4)"
    return rms


def test_single_speaker_cfg2_shape():
    from tts_b200.vits import VitsConfig
    got, want = _run(VitsConfig(), b=4, t=24, seed=11)
    rms = _check(got, want)
    b, _, n = got["model_outputs"].shape
    assert n == got["y_mask"].shape[-1] * 256  # tests/tts_tests/test_vits.py:283-290 shape contract
    print("waveform rms err", rms)


def test_ragged_lengths_and_length_scale():
    from tts_b200.vits import VitsConfig
    got, want = _run(VitsConfig(), b=3, t=30, seed=12, lengths=torch.tensor([30, 17, 4]), length_scale=1.7)
    _check(got, want)


def test_multispeaker_cfg5_shape():
    from tts_b200.vits import VitsArgs, VitsConfig
    cfg = VitsConfig(model_args=VitsArgs(use_speaker_embedding=True, num_speakers=109))
    got, want = _run(cfg, b=3, t=20, seed=13, lengths=torch.tensor([20, 11, 7]), speaker_ids=torch.tensor([3, 108, 0]))
    _check(got, want)
    from tts_b200.vits import Vits
    with pytest.raises(ValueError):
        Vits(cfg).cuda().inference(torch.zeros(1, 4, dtype=torch.long).cuda(), {"x_lengths": torch.tensor([4]).cuda()})
