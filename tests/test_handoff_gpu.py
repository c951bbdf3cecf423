"""The device hand-off kernels around the vocoder and the product model against fixtures produced by the REAL reference
classes (tests/golden/make_golden.py: AudioProcessor, interpolate_vocoder_input, the Synthesizer chain, save_wav's
arithmetic, Vits.inference itself)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _norm(kw):
    from tts_b200.vocoder import AudioNorm
    return AudioNorm(**{k: v for k, v in kw.items()})


def test_audio_processor_normalize_denormalize_all_branches(golden):
    gd = golden("vocoder_handoff")
    for case in gd["normalize_cases"]:
        ap = _norm(case["kw"])
        n = ap.normalize(case["S"].cuda()).cpu()
        assert torch.equal(n, case["normalized"]), case["kw"]
        d = ap.denormalize(case["denorm_input"].cuda()).cpu()
        assert torch.equal(d, case["denormalized"]), case["kw"]
    # batched [B,C,T] and strided [B,T,C] inputs address the same arithmetic
    case = gd["normalize_cases"][0]
    from tts_b200.vocoder import AudioNorm, vocoder_input
    s = case["S"].cuda()
    both = vocoder_input(torch.stack([s, s * 0.5]).transpose(1, 2).contiguous(), AudioNorm.identity(), _norm(case["kw"]),
                         time_last=False)
    assert torch.equal(both[0].cpu(), case["normalized"])


def test_interpolate_vocoder_input_vs_reference(golden):
    from tts_b200.vocoder import interpolate_vocoder_input
    for case in golden("vocoder_handoff")["interpolate_cases"]:
        got = interpolate_vocoder_input([1, case["scale"]], case["spec"].cuda()).cpu()
        assert got.shape == case["out"].shape, (case["scale"], got.shape, case["out"].shape)
        assert (got - case["out"]).abs().max() <= 2e-6 * max(1.0, float(case["out"].abs().max())), case["scale"]


def test_synthesizer_handoff_chain_in_one_pass(golden):
    """denormalize (TTS AP) -> normalize (vocoder AP) -> interpolate 22.05k -> 24k -> replicate pad 5  (synthesizer.py:412-429,
    hifigan_generator.py:281) from the TTS model's [T, C] output, one kernel."""
    from tts_b200.vocoder import vocoder_input
    ch = golden("vocoder_handoff")["chain"]
    got = vocoder_input(ch["mel_tc"].cuda().unsqueeze(0), _norm(ch["tts_kw"]), _norm(ch["voc_kw"]),
                        scale_factor=ch["sr_voc"] / ch["sr_tts"], padding=5, time_last=False).cpu()
    assert got.shape == ch["out"].shape
    assert (got - ch["out"]).abs().max() <= 5e-6 * max(1.0, float(ch["out"].abs().max()))


def test_save_wav_int16_on_device(golden):
    from tts_b200.vocoder import new_peak, wav_to_int16
    gd = golden("vocoder_handoff")
    got = wav_to_int16(gd["wav"].cuda()).cpu()
    assert torch.equal(got, gd["wav_int16"])
    assert wav_to_int16(torch.zeros(8).cuda()).tolist() == [0] * 8                 # max(0.01, peak) guard
    p = new_peak("cuda")
    wav_to_int16(gd["wav"].cuda() * 0.5)                                          # a peak word can be shared / pre-folded
    assert int(p.item()) == 0


def test_gan_inference_and_config_surface():
    import vits_oracle as O
    from tts_b200.vocoder import GAN, HifiganConfig
    torch.manual_seed(2)
    cfg = HifiganConfig()
    cfg.generator_model_params["upsample_initial_channel"] = 64            # narrow: keeps the CPU oracle fast
    gan = GAN(cfg).eval()
    mel = torch.randn(1, 80, 21)
    want = O.hifigan_inference(gan.model_g.state_dict(), mel, 5)
    gan.cuda()
    got = gan.inference(mel.cuda()).cpu()
    assert got.shape == want.shape == (1, 1, (21 + 10) * 256)
    assert (got - want).pow(2).mean().sqrt() <= 1e-5


def test_product_matches_the_real_reference_model_fixture(golden):
    """tests/golden/vits_real_model_small.pt was written by TTS.tts.models.vits.Vits.inference (multi-speaker)."""
    from tts_b200.vits import Vits, VitsArgs, VitsConfig
    g = golden("vits_real_model_small")
    fields = {k: v for k, v in g["args"].items() if k in VitsArgs.__dataclass_fields__}
    m = Vits(VitsConfig(model_args=VitsArgs(**fields))).eval()
    missing, unexpected = m.load_state_dict(g["state"], strict=False)
    assert not unexpected and all(k.startswith(("posterior_encoder.", "duration_predictor.post_")) for k in missing)
    m.cuda()
    out = m.inference(g["tokens"].cuda(), {"x_lengths": g["x_lengths"].cuda(), "speaker_ids": g["speaker_ids"].cuda()},
                      sdp_noise=g["sdp_noise"], prior_noise=g["prior_noise"].cuda())
    want = g["out"]
    assert torch.equal(out["durations"].cpu(), want["durations"])
    assert torch.equal(out["alignments"].cpu(), want["alignments"])
    assert torch.equal(out["y_mask"].cpu(), want["y_mask"])
    for k in ("m_p", "logs_p", "z_p", "z"):
        assert (out[k].cpu() - want[k]).abs().max() < 2e-4, k
    err = out["model_outputs"].cpu() - want["model_outputs"]
    rms, ref = float(err.pow(2).mean().sqrt()), float(want["model_outputs"].pow(2).mean().sqrt())
    assert rms <= 1e-4 and rms <= 1e-4 * ref, (rms, ref)


def test_synthesize_batched_on_device_vs_sentence_at_a_time_oracle():
    """what Synthesizer.tts does one sentence per call (synthesizer.py:384-441), as three length buckets
    on the device.  Per sentence: durations / alignment bit-exact vs the oracle's batch-1 inference; waveform equal up
    to the decoder's receptive field before the sentence's end (there a padded row differs from a solo run in the
    reference too); then the joined int16 stream equals save_wav's arithmetic on the same float samples."""
    from dataclasses import asdict

    import numpy as np

    import vits_oracle as O
    from tts_b200.parallel import concat_sentences, synthesize_batched, synthesize_to_int16
    from tts_b200.vits import Vits, VitsArgs, VitsConfig
    torch.manual_seed(31)
    args = VitsArgs(upsample_initial_channel_decoder=64, num_layers_text_encoder=2, hidden_channels_ffn_text_encoder=256)
    m = Vits(VitsConfig(model_args=args)).eval()
    gen = torch.Generator().manual_seed(32)
    for _, p in m.named_parameters():
        if float(p.detach().abs().sum()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=gen) * 0.05)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.cuda()
    lens = [23, 7, 40, 22, 8, 39, 3]
    seqs = [torch.randint(1, 100, (n,), generator=gen).tolist() for n in lens]
    sdp = [torch.randn(2, n, generator=gen) for n in lens]
    prior = lambda i, c, t: torch.randn(c, t, generator=torch.Generator().manual_seed(1000 + i))
    got = synthesize_batched(m, seqs, max_padded_tokens=96, max_batch=3, sdp_noise=sdp, prior_noise=prior)
    from tts_b200.parallel import bucket_by_length
    assert len(bucket_by_length(lens, 96, 3)) >= 3
    a = asdict(args)
    for i, s in enumerate(seqs):
        want = O.vits_inference(sd, torch.tensor([s]), torch.tensor([len(s)]), sdp[i].unsqueeze(0),
                                lambda shape, i=i: prior(i, shape[1], shape[2]).unsqueeze(0), args=a)
        n = int(want["y_lengths"][0]) * 256
        assert got[i].shape == (n,), (i, got[i].shape, n)
        keep = max(0, n - 16 * 256)                         # 16 frames > the decoder's reach past a sentence's end
        err = got[i][:keep].cpu() - want["model_outputs"][0, 0, :keep]
        ref = want["model_outputs"][0, 0, :keep]
        if keep:
            assert float(err.pow(2).mean().sqrt()) <= 1e-4 * max(float(ref.pow(2).mean().sqrt()), 1e-6), i
    joined = concat_sentences(got, gap=10000)
    want16 = O.wav_to_int16(joined.cpu().numpy())
    got16 = synthesize_to_int16(m, seqs, max_padded_tokens=96, max_batch=3, sdp_noise=sdp, prior_noise=prior)
    assert got16.dtype == torch.int16 and np.array_equal(got16.cpu().numpy(), want16)
