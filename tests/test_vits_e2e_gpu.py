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
    ref = want["model_outputs"].pow(2).mean().sqrt().item()
    # random-init audio is quiet (RMS ~0.03): the absolute bound alone would also pass a single-pass-TF32 regression
    assert rms <= 1e-4 * ref, f"relative waveform RMS error {rms / ref}"
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


def test_deterministic_duration_predictor_model():
    """VitsArgs.use_sdp=False (vits.py:646-654, :1137-1139)."""
    from tts_b200.vits import VitsArgs, VitsConfig
    cfg = VitsConfig(model_args=VitsArgs(use_sdp=False, use_speaker_embedding=True, num_speakers=10))
    got, want = _run(cfg, b=3, t=22, seed=14, lengths=torch.tensor([22, 13, 6]), speaker_ids=torch.tensor([1, 9, 0]))
    _check(got, want)


def test_external_durations():
    """aux_input['durations'] replaces the duration predictor (vits.py:1141-1143; single utterance)."""
    from tts_b200.vits import Vits, VitsConfig
    torch.manual_seed(15)
    cfg = VitsConfig()
    m = Vits(cfg).eval()
    _perturb(m, 15)
    tok = torch.randint(0, 100, (1, 19))
    dur = torch.randint(0, 5, (1, 19)).float() + torch.rand(1, 19) * 0.5
    sd = m.state_dict()
    # oracle: same glue with w = durations.unsqueeze(0)
    x, m_p, logs_p, x_mask = O.text_encoder(O.sub(sd, "text_encoder"), tok, torch.tensor([19]))
    w_ceil = torch.ceil(dur.unsqueeze(0))
    y_len = torch.clamp_min(w_ceil.sum([1, 2]), 1).long()
    y_mask = O.sequence_mask(y_len, None).float().unsqueeze(1)
    attn = O.generate_path(w_ceil.squeeze(1), (x_mask * y_mask.transpose(1, 2)).squeeze(1).transpose(1, 2))
    noise = torch.randn(1, 192, int(y_len))
    mp = torch.matmul(attn.transpose(1, 2), m_p.transpose(1, 2)).transpose(1, 2)
    lp = torch.matmul(attn.transpose(1, 2), logs_p.transpose(1, 2)).transpose(1, 2)
    z_p = mp + noise * torch.exp(lp) * 0.667
    z = O.flow_forward(O.sub(sd, "flow"), z_p, y_mask, reverse=True)
    wav = O.hifigan_forward(O.sub(sd, "waveform_decoder"), z * y_mask)
    m.cuda()
    got = m.inference(tok.cuda(), {"x_lengths": torch.tensor([19]).cuda(), "durations": dur.cuda()},
                      prior_noise=noise.cuda())
    assert torch.equal(got["durations"].cpu(), w_ceil)
    assert torch.equal(got["alignments"].cpu(), attn)
    assert (got["z_p"].cpu() - z_p).abs().max() < 2e-4
    rms = (got["model_outputs"].cpu() - wav).pow(2).mean().sqrt().item()
    assert rms <= 1e-4, rms
    with pytest.raises(ValueError):
        m.inference(torch.cat([tok, tok]).cuda(), {"durations": dur.cuda()})


@pytest.mark.parametrize("mode", ["speaker_embedding", "d_vector"])
def test_voice_conversion(mode):
    """Vits.voice_conversion / inference_voice_conversion (vits.py:1175-1232) vs the oracle's restated glue."""
    import torch.nn.functional as F
    from tts_b200.audio import wav_to_spec
    from tts_b200.vits import Vits, VitsArgs, VitsConfig
    torch.manual_seed(16)
    if mode == "speaker_embedding":
        cfg = VitsConfig(model_args=VitsArgs(use_speaker_embedding=True, num_speakers=7))
    else:
        cfg = VitsConfig(model_args=VitsArgs(use_d_vector_file=True, d_vector_dim=64, num_speakers=7))
    m = Vits(cfg).eval()
    _perturb(m, 16)
    sd = m.state_dict()
    wav = (torch.rand(2, 1, 256 * 45) * 2 - 1) * 0.7
    if mode == "speaker_embedding":
        src, tgt = torch.tensor([2, 5]), torch.tensor([6, 0])
        g_src, g_tgt = (F.embedding(i, sd["emb_g.weight"]).unsqueeze(-1) for i in (src, tgt))
    else:
        src, tgt = torch.randn(2, 64), torch.randn(2, 64)
        g_src, g_tgt = F.normalize(src).unsqueeze(-1), F.normalize(tgt).unsqueeze(-1)
    y = O.wav_to_spec(wav, 1024, 256, 1024)
    lens = torch.tensor([y.shape[-1], 20])
    noise = torch.randn(2, 192, y.shape[-1])
    want_o, want_mask, (wz, wzp, wzh) = O.voice_conversion(sd, y, lens, g_src, g_tgt, noise, args=_args_dict(cfg.model_args))
    m.cuda()
    y_dev = wav_to_spec(wav.cuda(), 1024, 256, 1024)
    assert (y_dev.cpu() - y).abs().max() < 2e-3 * max(1.0, y.abs().max().item())
    o, mask, (z, zp, zh) = m.voice_conversion(y.cuda(), lens.cuda(), src.cuda(), tgt.cuda(), posterior_noise=noise.cuda())
    assert torch.equal(mask.cpu(), want_mask)
    for a, w, n in ((z, wz, "z"), (zp, wzp, "z_p"), (zh, wzh, "z_hat")):
        err = ((a.cpu() - w) * want_mask).abs().max().item()
        assert err < 5e-4, (n, err)
    assert o.shape == want_o.shape
    rms = (o.cpu() - want_o).pow(2).mean().sqrt().item()
    assert rms <= 1e-4, rms
    # the wav-in / wav-out entry point: same shapes, finite, and equal to voice_conversion on its own spectrogram
    kw = dict(speaker_id=tgt.cuda(), reference_speaker_id=src.cuda()) if mode == "speaker_embedding" else \
        dict(d_vector=tgt.cuda(), reference_d_vector=src.cuda())
    o2 = m.inference_voice_conversion(wav.cuda(), posterior_noise=noise.cuda(), **kw)
    o3, _, _ = m.voice_conversion(y_dev, torch.tensor([y.shape[-1]] * 2).cuda(), src.cuda(), tgt.cuda(),
                                  posterior_noise=noise.cuda())
    assert torch.equal(o2, o3) and torch.isfinite(o2).all() and o2.shape == (2, 1, y.shape[-1] * 256)


def test_encoder_sample_rate_upsampling():
    """VitsArgs.encoder_sample_rate: z is linearly interpolated by sample_rate / encoder_sample_rate before the
    decoder and the mask is rebuilt (Vits.upsampling_z, vits.py:944-959)."""
    from tts_b200.vits import Vits, VitsArgs, VitsAudioConfig, VitsConfig
    args = VitsArgs(encoder_sample_rate=11025, upsample_rates_decoder=[8, 8, 4, 2],
                    upsample_kernel_sizes_decoder=[16, 16, 8, 4])
    cfg = VitsConfig(model_args=args, audio=VitsAudioConfig(sample_rate=22050))
    assert Vits.init_from_config(cfg) is not None          # 8*8*4*2 == 256 * 2 (vits.py:1789-1794)
    with pytest.raises(AssertionError):
        Vits.init_from_config(VitsConfig(model_args=VitsArgs(encoder_sample_rate=11025)))
    torch.manual_seed(21)
    m = Vits(cfg).eval()
    _perturb(m, 21)
    a = _args_dict(args)
    a["sample_rate"] = 22050
    b, t = 3, 18
    tok, lens = torch.randint(0, 100, (b, t)), torch.tensor([18, 9, 3])
    sdp_noise = torch.randn(b, 2, t)
    store = {}

    def prior_noise(shape):
        store["n"] = torch.randn(shape, generator=torch.Generator().manual_seed(5))
        return store["n"]

    want = O.vits_inference(m.state_dict(), tok, lens, sdp_noise, prior_noise, args=a)
    m.cuda()
    got = m.inference(tok.cuda(), {"x_lengths": lens.cuda()}, sdp_noise=sdp_noise,
                      prior_noise=lambda s: store["n"].cuda())
    assert got["z"].shape == want["z"].shape and got["z"].shape[-1] == 2 * got["z_p"].shape[-1]
    assert torch.equal(got["y_mask"].cpu(), want["y_mask"])
    assert (got["z"].cpu() - want["z"]).abs().max() < 2e-4
    assert got["model_outputs"].shape == want["model_outputs"].shape == (b, 1, got["z"].shape[-1] * 512)
    rms = (got["model_outputs"].cpu() - want["model_outputs"]).pow(2).mean().sqrt().item()
    assert rms <= 1e-4, rms


def test_upsample_linear_matches_torch():
    import torch.nn.functional as F
    from tts_b200.layers import upsample_linear
    torch.manual_seed(3)
    for f in (2.0, 1.5, 22050 / 16000):
        z = torch.randn(2, 7, 40)
        want = F.interpolate(z, scale_factor=[f], mode="linear")
        got = upsample_linear(z.cuda(), f).cpu()
        assert got.shape == want.shape
        assert (got - want).abs().max() < 1e-6, f
