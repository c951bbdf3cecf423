"""Parity in the regime the benchmark runs in.

The tcgen05 conv kernels are persistent: one CTA per SM loops over (batch, row tile, time tile) with cross-tile state
(double-buffered TMEM accumulator phase, operand ring parities).  The small-shape tests never give a CTA more than one
tile; these do -- every case below launches 10..45 tiles per CTA, the bench's regime -- and they run the BASELINE
configurations at their real sizes against the CPU oracle.  Where the oracle would need minutes for the full batch it
runs on a subset of rows that contains the longest utterance (rows never interact on the path, and with the longest row
present the padded length -- hence every row's arithmetic -- is the same as in the full batch).

Tolerances: integer-valued outputs bit-exact; waveforms abs RMS <= 1e-4 (north_star) AND relative RMS <= 1e-4
(a single-pass-TF32 regression sits at ~1e-3 relative and fails); single layers relative RMS <= 5e-5.
"""
import os
from dataclasses import asdict

import pytest
import torch
import torch.nn.functional as F

import vits_oracle as O

pytestmark = pytest.mark.gpu
FULL = bool(int(os.environ.get("B200TTS_FULL_TESTS", "0")))
# one 3xTF32 layer on unit-variance data: the tensor core truncates when it accumulates into fp32 TMEM, which grows with
# the reduction length Cin*K (measured 1.0e-5 at 128x11, 2.0e-5 at 256x11); single-pass TF32 sits at ~3e-4
LAYER_REL_TOL = 5e-5


def _rel_rms(got, want):
    err = (got.double() - want.double())
    return float(err.pow(2).mean().sqrt() / want.double().pow(2).mean().sqrt().clamp_min(1e-30)), float(err.abs().max())


# ----------------------------------------------------------------------------- single layers, many tiles per CTA
LAYER_CASES = [
    # (C, K, dil, B, T, expected kernel family)        tiles = B * ceil(T / 256 or 240) * row tiles  (148 CTAs)
    (128, 11, 5, 32, 9600, "tc3"),          # 1216 x 1 tiles: stage-1 MRF, the FLOP carrier
    (128, 3, 1, 32, 9600, None),            # tc3 or tc3_staged (K <= 3 wide layer)
    (128, 7, 3, 32, 9600, "tc3"),
    (256, 7, 1, 32, 2400, "tc3"),           # 2 row tiles
    (256, 11, 5, 16, 4800, "tc3"),
    (64, 11, 1, 32, 19200, "tc3_grouped"),  # GRP = 2
    (64, 3, 3, 32, 19200, "tc3_grouped"),
    (64, 7, 5, 32, 19200, "tc3_grouped"),
    (32, 7, 5, 32, 38400, "tc3_grouped"),   # GRP = 4
    (32, 3, 1, 32, 38400, "tc3_grouped"),
    (32, 11, 3, 32, 38400, "tc3_grouped"),
]


@pytest.mark.parametrize("c,k,dil,b,t,family", LAYER_CASES)
def test_conv_layer_many_tiles_per_cta(c, k, dil, b, t, family):
    """ResBlock1's second-conv form: y_old + ((conv(lrelu(x)) + bias) + residual), MRF mean on top."""
    from tts_b200 import _lib
    from tts_b200.conv import FusedConv1d
    torch.manual_seed(c * 1000 + k * 10 + dil)
    w = torch.randn(c, c, k) / (c * k) ** 0.5
    bias = torch.randn(c) * 0.1
    x = torch.randn(b, c, t)
    res = torch.randn(b, c, t)
    yold = torch.randn(b, c, t)
    pad = (k * dil - dil) // 2
    want = (yold + (F.conv1d(F.leaky_relu(x, 0.1), w, bias, dilation=dil, padding=pad) + res)) / 3.0
    conv = FusedConv1d(w, bias, dilation=dil, padding=pad)
    y = yold.cuda().clone()
    with _lib.dispatch_log() as log:
        got = conv(x.cuda(), in_slope=0.1, residual=res.cuda(), accumulate_into=y, post_div=3.0)
    torch.cuda.synchronize()
    assert _lib.lib().b200tts_debug_tc_error() == 0
    if family is not None:
        assert log.names == [family], log.names
    else:
        assert log.names in (["tc3"], ["tc3_staged"]), log.names
    rel, mx = _rel_rms(got.cpu(), want)
    assert rel <= LAYER_REL_TOL and mx <= 2e-4 * float(want.abs().max()), (rel, mx)
    # plain form (no residual / accumulate), a different tile count through the same persistent loop
    got2 = conv(x[:, :, : t - 77].cuda(), in_slope=0.1)
    want2 = F.conv1d(F.leaky_relu(x[:, :, : t - 77], 0.1), w, bias, dilation=dil, padding=pad)
    rel2, _ = _rel_rms(got2.cpu(), want2)
    assert rel2 <= LAYER_REL_TOL, rel2


@pytest.mark.parametrize("cin,cout,k,s,b,t", [(256, 128, 16, 8, 32, 1200), (512, 256, 16, 8, 32, 152),
                                              (128, 64, 4, 2, 32, 9600), (64, 32, 4, 2, 32, 19200)])
def test_upsampler_many_tiles_per_cta(cin, cout, k, s, b, t):
    """o = ups(leaky_relu(o, 0.1))  (hifigan_generator.py:248-249) as a polyphase conv on the tcgen05 kernel."""
    from tts_b200 import _lib
    from tts_b200.conv import FusedConv1d
    torch.manual_seed(cin + k)
    w = torch.randn(cin, cout, k) / (cin * k / s) ** 0.5
    bias = torch.randn(cout) * 0.1
    x = torch.randn(b, cin, t)
    want = F.conv_transpose1d(F.leaky_relu(x, 0.1), w, bias, stride=s, padding=(k - s) // 2)
    conv = FusedConv1d(w, bias, padding=(k - s) // 2, transposed=True, stride=s)
    with _lib.dispatch_log() as log:
        got = conv(x.cuda(), in_slope=0.1)
    torch.cuda.synchronize()
    assert _lib.lib().b200tts_debug_tc_error() == 0
    assert log.names == ["tc3"], log.names
    assert got.shape == want.shape
    rel, mx = _rel_rms(got.cpu(), want)
    assert rel <= LAYER_REL_TOL and mx <= 2e-4 * float(want.abs().max()), (rel, mx)


# ----------------------------------------------------------------------------- which kernel each layer takes
def test_decoder_and_flow_dispatch_is_pinned():
    """A dispatch change must not silently move the hot path onto a fallback generation (VERDICT weak #10)."""
    from tts_b200 import _lib
    from tts_b200.vits import Vits, VitsConfig
    torch.manual_seed(0)
    m = Vits(VitsConfig()).eval().cuda()
    z = torch.randn(2, 192, 256).cuda()
    with _lib.dispatch_log() as log:
        m.waveform_decoder(z)
    names = log.names
    # conv_pre, then per stage: ups + 18 resblock convs, then conv_post
    assert names[0] == "tc3" and names[-1] == "row1", names
    body = names[1:-1]
    fused = "resblock" in body
    if not fused:
        assert len(body) == 4 * 19, len(body)
        for s in range(4):
            stage = body[s * 19:(s + 1) * 19]
            assert stage[0] == "tc3", (s, stage)
            want = {"tc3", "tc3_staged"} if s < 2 else {"tc3_grouped"}
            assert set(stage[1:]) <= want, (s, stage)
    assert not ({"tc1", "tc2"} & set(names)), names          # the superseded generations are never on the bench path
    with _lib.dispatch_log() as log:                         # a frame count that is not a multiple of 4 (unaligned rows)
        m.waveform_decoder(torch.randn(2, 192, 150).cuda())
    assert log.names[0] == "tc3" and log.names[1] == "tc3" and "fma" not in log.names, log.names
    mask = torch.ones(4, 1, 192).cuda()
    with _lib.dispatch_log() as log:
        m.flow(torch.randn(4, 192, 192).cuda(), mask, reverse=True)
    assert set(log.names) <= {"tc3", "tc3_staged"} and len(log.names) == 4 * (2 + 2 * 4), log.names


# ----------------------------------------------------------------------------- decoder at length
def _decoder_args(cin=192, cond=0):
    return dict(in_channels=cin, out_channels=1, resblock_type="1", resblock_dilation_sizes=[[1, 3, 5]] * 3,
                resblock_kernel_sizes=[3, 7, 11], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
                upsample_factors=[8, 8, 2, 2], inference_padding=0, cond_channels=cond, conv_pre_weight_norm=False,
                conv_post_weight_norm=False, conv_post_bias=False)


def _wav_check(got, want, valid=None):
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape
    err = got - want
    if valid is not None:
        err, want = err * valid, want * valid
        n = float(valid.sum())
    else:
        n = float(err.numel())
    rms = float((err.pow(2).sum() / n).sqrt())
    ref = float((want.pow(2).sum() / n).sqrt())
    assert rms <= 1e-4, f"waveform RMS error {rms} (north_star bound 1e-4)"
    assert rms <= 1e-4 * ref, f"relative waveform RMS error {rms / ref} (signal RMS {ref})"
    return rms, ref


def test_full_decoder_b4_t1024_vs_oracle():
    """4 x 1024 frames: 1024..16384 tiles per launch through the persistent kernels (cfg3's length)."""
    from tts_b200.hifigan import HifiganGenerator
    torch.manual_seed(41)
    a = _decoder_args()
    m = HifiganGenerator(**a).eval()
    x = torch.randn(4, 192, 1024)
    want = O.hifigan_forward(m.state_dict(), x)
    got = m.cuda()(x.cuda())
    assert got.shape == (4, 1, 1024 * 256)
    rms, ref = _wav_check(got, want)
    print(f"decoder B=4 T=1024: rms err {rms:.3e}, signal rms {ref:.3e}")


def test_cfg1_standalone_hifigan_4x80x256():
    """BASELINE configs[0]: HifiganGenerator(80, 1, '1', ...) on randn(4, 80, 256), weight norm removed -> [4,1,65536]."""
    from tts_b200.hifigan import HifiganGenerator
    torch.manual_seed(1234)
    m = HifiganGenerator(80, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2]).eval()
    mel = torch.randn(4, 80, 256)
    want = O.hifigan_forward(m.state_dict(), mel)
    m.remove_weight_norm()
    got = m.cuda()(mel.cuda())
    assert got.shape == (4, 1, 65536)
    _wav_check(got, want)


def test_cfg3_shard_flow_then_hifigan_t1024():
    """BASELINE configs[2] per-GPU shard: z_p [b,192,1024], mask = 1 -> flow reverse -> HiFiGAN(192)."""
    from tts_b200.vits import Vits, VitsConfig
    b = 32 if FULL else 6
    torch.manual_seed(1234)
    m = Vits(VitsConfig()).eval()
    gen = torch.Generator().manual_seed(3)
    for _, p in m.flow.named_parameters():       # the reference zero-initialises `post`: perturb so the flow is not vacuous
        if float(p.abs().sum()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=gen) * 0.05)
    sd = m.state_dict()
    z_p = torch.randn(b, 192, 1024)
    mask = torch.ones(b, 1, 1024)
    want_z = O.flow_forward(O.sub(sd, "flow"), z_p, mask, reverse=True)
    want = O.hifigan_forward(O.sub(sd, "waveform_decoder"), want_z * mask)
    m.cuda()
    z = m.flow(z_p.cuda(), mask.cuda(), reverse=True)
    rel, mx = _rel_rms(z.cpu(), want_z)
    assert rel <= 2e-5 and mx <= 2e-4, (rel, mx)
    got = m.waveform_decoder(z * mask.cuda())
    assert got.shape == (b, 1, 262144)
    _wav_check(got, want)


# ----------------------------------------------------------------------------- end to end at the BASELINE sizes
def _perturb(m, seed):
    gen = torch.Generator().manual_seed(seed)
    for _, p in m.named_parameters():
        if float(p.abs().sum()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=gen) * 0.05)


def _e2e(cfg, tok, lens, seed, length_scale=1.0, speaker_ids=None, oracle_rows=None):
    from tts_b200.vits import Vits
    torch.manual_seed(seed)
    m = Vits(cfg).eval()
    _perturb(m, seed)
    m.length_scale = length_scale
    a = asdict(cfg.model_args)
    a["length_scale"] = length_scale
    b, t = tok.shape
    sdp_noise = torch.randn(b, 2, t, generator=torch.Generator().manual_seed(seed + 1))
    store = {}

    def prior_noise(shape):
        store["n"] = torch.randn(shape, generator=torch.Generator().manual_seed(seed + 2))
        return store["n"].cuda()

    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.cuda()
    aux = {"x_lengths": lens.cuda(), "speaker_ids": None if speaker_ids is None else speaker_ids.cuda()}
    got = m.inference(tok.cuda(), aux, sdp_noise=sdp_noise, prior_noise=prior_noise)
    torch.cuda.synchronize()
    ylen = got["y_lengths"].cpu()
    if oracle_rows is None:
        rows = torch.arange(b)
    else:   # a subset that contains the longest utterance: same padded length, hence same arithmetic per row
        longest = int(ylen.argmax())
        others = [i for i in torch.randperm(b, generator=torch.Generator().manual_seed(seed + 3)).tolist() if i != longest]
        rows = torch.tensor(sorted([longest] + others[: oracle_rows - 1]))
    want = O.vits_inference(sd, tok[rows], lens[rows], sdp_noise[rows], lambda s: store["n"][rows], args=a,
                            speaker_ids=None if speaker_ids is None else speaker_ids[rows])
    assert torch.equal(got["durations"].cpu()[rows], want["durations"]), "durations differ"
    assert torch.equal(ylen[rows], want["y_lengths"])
    assert torch.equal(got["alignments"].cpu()[rows], want["alignments"]), "alignment path differs"
    assert torch.equal(got["y_mask"].cpu()[rows], want["y_mask"])
    assert torch.equal(got["wav_lengths"].cpu()[rows], want["y_lengths"] * 256)
    for k in ("m_p", "logs_p", "z_p", "z"):
        err = (got[k].cpu()[rows] - want[k]).abs().max().item()
        assert err < 2e-4, (k, err)
    n = want["model_outputs"].shape[-1]
    valid = (torch.arange(n)[None, None, :] < (want["y_lengths"] * 256)[:, None, None]).float()
    rms, ref = _wav_check(got["model_outputs"].cpu()[rows], want["model_outputs"], valid)
    _wav_check(got["model_outputs"].cpu()[rows], want["model_outputs"])      # the padded tail as well
    return rms, ref, got


def test_cfg2_b32_t64_length_scale_1():
    """BASELINE configs[1] exactly as bench.py runs it: 32 utterances x 64 tokens, VitsArgs() defaults."""
    from tts_b200.vits import VitsConfig
    gen = torch.Generator().manual_seed(4321)
    tok = torch.randint(0, 100, (32, 64), generator=gen)
    rms, ref, got = _e2e(VitsConfig(), tok, torch.full((32,), 64), seed=1234)
    print(f"cfg2: frames {got['y_mask'].shape[-1]}, wav rms err {rms:.3e} (signal {ref:.3e})")


def test_cfg2_b32_t64_length_scale_3():
    """Same batch at length_scale 3 (~450 frames = ~5 s per utterance, LJSpeech-shaped work)."""
    from tts_b200.vits import VitsConfig
    gen = torch.Generator().manual_seed(4321)
    tok = torch.randint(0, 100, (32, 64), generator=gen)
    _e2e(VitsConfig(), tok, torch.full((32,), 64), seed=1234, length_scale=3.0, oracle_rows=None if FULL else 6)


def test_cfg5_multispeaker_b128_mixed_lengths():
    """BASELINE configs[4]: 109 speakers, batch 128, x_lengths ~ U[20,128], padded tokens, length masking."""
    from tts_b200.vits import VitsArgs, VitsConfig
    cfg = VitsConfig(model_args=VitsArgs(use_speaker_embedding=True, num_speakers=109))
    gen = torch.Generator().manual_seed(55)
    lens = torch.randint(20, 129, (128,), generator=gen)
    lens[7] = 128
    tok = torch.randint(0, 100, (128, 128), generator=gen) * (torch.arange(128)[None, :] < lens[:, None])
    sid = torch.randint(0, 109, (128,), generator=gen)
    _e2e(cfg, tok, lens, seed=77, speaker_ids=sid, oracle_rows=None if FULL else 8)


def test_two_checkpoints_in_a_row_rebuild_the_decoder(tmp_path):
    """Vits.load_checkpoint twice must not keep the first checkpoint's packed decoder weights."""
    from tts_b200.vits import Vits, VitsConfig
    cfg = VitsConfig()
    paths = []
    for seed in (1, 2):
        torch.manual_seed(seed)
        src = Vits(cfg).eval()
        p = tmp_path / f"ck{seed}.pth"
        torch.save({"model": src.state_dict()}, p)
        paths.append((p, {k: v.clone() for k, v in src.state_dict().items()}))
    m = Vits(cfg).eval().cuda()
    z = torch.randn(1, 192, 12)
    outs = []
    for p, sd in paths:
        m.load_checkpoint(cfg, str(p), eval=True)
        got = m.waveform_decoder(z.cuda()).cpu()
        want = O.hifigan_forward(O.sub(sd, "waveform_decoder"), z)
        assert (got - want).pow(2).mean().sqrt() < 1e-5
        outs.append(got)
    assert (outs[0] - outs[1]).abs().max() > 1e-4
