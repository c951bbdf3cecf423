"""CUDA text encoder / SDP / flow / path expansion vs golden fixtures (reference outputs) and the oracle."""
import pytest
import torch

import vits_oracle as O

pytestmark = pytest.mark.gpu


def _close(got, want, atol, name=""):
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape, (name, got.shape, want.shape)
    err = (got - want).abs().max().item()
    assert err <= atol, (name, err, want.abs().max().item())


def _perturb(m):
    for _, p in m.named_parameters():
        if float(p.abs().sum()) == 0.0:
            p.data.normal_(0, 0.05)


def test_text_encoder_golden(golden):
    from tts_b200.layers import TextEncoder
    g = golden("text_encoder_small")
    m = TextEncoder(**g["args"]).eval()
    m.load_state_dict(g["state"])
    m.cuda()
    x, mp, logs, mask = m(g["tokens"].cuda(), g["lengths"].cuda())
    _close(x, g["x"], 2e-5, "x")
    _close(mp, g["m_p"], 2e-5, "m_p")
    _close(logs, g["logs_p"], 2e-5, "logs_p")
    assert torch.equal(mask.cpu(), g["x_mask"])


@pytest.mark.parametrize("b,t", [(3, 21), (2, 64), (1, 5), (2, 130)])
def test_text_encoder_full_width_vs_oracle(b, t):
    from tts_b200.layers import TextEncoder
    torch.manual_seed(t)
    m = TextEncoder(100, 192, 192, 768, 2, 6, 3, 0.1).eval()
    tok = torch.randint(0, 100, (b, t))
    lens = torch.randint(1, t + 1, (b,))
    lens[0] = t
    want = O.text_encoder(m.state_dict(), tok, lens)
    m.cuda()
    got = m(tok.cuda(), lens.cuda())
    for a, w, n in zip(got, want, ("x", "m", "logs", "mask")):
        _close(a, w, 5e-5, n)


def test_text_encoder_language_embedding_vs_oracle():
    from tts_b200.layers import TextEncoder
    torch.manual_seed(2)
    m = TextEncoder(50, 192, 192, 768, 2, 2, 3, 0.1, language_emb_dim=4).eval()
    tok, lens = torch.randint(0, 50, (2, 17)), torch.tensor([17, 9])
    le = torch.randn(2, 4, 1)
    want = O.text_encoder(m.state_dict(), tok, lens, le, hidden=192, out_channels=192, num_layers=2)
    m.cuda()
    got = m(tok.cuda(), lens.cuda(), lang_emb=le.cuda())
    for a, w, n in zip(got, want, ("x", "m", "logs", "mask")):
        _close(a, w, 5e-5, n)


def test_flow_golden(golden):
    from tts_b200.layers import ResidualCouplingBlocks
    g = golden("flow_small")
    m = ResidualCouplingBlocks(**g["args"]).eval()
    m.load_state_dict(g["state"])
    m.cuda()
    got = m(g["z"].cuda(), g["mask"].cuda(), g=g["g"].cuda(), reverse=True)
    _close(got, g["rev"], 2e-5, "flow reverse")
    got = m(g["z"].cuda(), g["mask"].cuda(), g=g["g"].cuda(), reverse=False)
    _close(got, g["fwd"], 2e-5, "flow forward")


@pytest.mark.parametrize("flows", [3, 4])
def test_flow_forward_then_reverse_is_identity(flows):
    """Size-independent property (networks.py:214-232): reverse(forward(z)) == z on the unmasked frames, also for an
    odd number of flows (one channel flip left over after folding the rest into the packed weights)."""
    from tts_b200.layers import ResidualCouplingBlocks
    torch.manual_seed(flows)
    m = ResidualCouplingBlocks(192, 192, 5, 1, 4, num_flows=flows, cond_channels=64).eval()
    _perturb(m)
    z, g = torch.randn(2, 192, 301), torch.randn(2, 64, 1)
    mask = (torch.arange(301)[None, :] < torch.tensor([301, 77])[:, None]).float().unsqueeze(1)
    kw = dict(num_flows=flows, hidden=192, kernel_size=5, dilation_rate=1, num_layers=4)
    want = O.flow_forward(m.state_dict(), z, mask, g, reverse=False, **kw)
    m.cuda()
    zp = m(z.cuda(), mask.cuda(), g=g.cuda(), reverse=False)
    _close(zp, want, 1e-4, "flow forward vs oracle")
    back = m(zp, mask.cuda(), g=g.cuda(), reverse=True)
    _close(back * mask.cuda(), z * mask, 2e-4, "round trip")


def test_posterior_encoder_golden(golden):
    from tts_b200.layers import PosteriorEncoder
    g = golden("posterior_small")
    m = PosteriorEncoder(**g["args"]).eval()
    m.load_state_dict(g["state"])
    m.cuda()
    z, mean, logs, mask = m(g["y"].cuda(), g["y_lengths"].cuda(), g=g["g"].cuda(), noise=g["noise"].cuda())
    _close(mean, g["mean"], 3e-5, "mean")
    _close(logs, g["log_scale"], 3e-5, "log_scale")
    _close(z, g["z"], 5e-5, "z")
    assert torch.equal(mask.cpu(), g["y_mask"])


def test_posterior_encoder_full_width_vs_oracle():
    from tts_b200.layers import PosteriorEncoder
    torch.manual_seed(11)
    m = PosteriorEncoder(513, 192, 192, 5, 1, 16, cond_channels=256).eval()
    y, g = torch.randn(2, 513, 203).abs(), torch.randn(2, 256, 1)
    lens = torch.tensor([203, 90])
    noise = torch.randn(2, 192, 203)
    want = O.posterior_encoder(m.state_dict(), y, lens, g=g, noise=noise)
    m.cuda()
    got = m(y.cuda(), lens.cuda(), g=g.cuda(), noise=noise.cuda())
    for a, w, n in zip(got, want, ("z", "mean", "log_scale", "mask")):
        _close(a, w, 2e-4, n)


def test_duration_predictor_golden(golden):
    from tts_b200.layers import DurationPredictor
    g = golden("duration_predictor_small")
    m = DurationPredictor(**g["args"]).eval()
    m.load_state_dict(g["state"])
    m.cuda()
    got = m(g["x"].cuda(), g["x_mask"].cuda(), g=g["g"].cuda(), lang_emb=g["lang_emb"].cuda())
    _close(got, g["logw"], 2e-5, "logw")


@pytest.mark.parametrize("cond", [0, 256])
def test_duration_predictor_full_width_vs_oracle(cond):
    from tts_b200.layers import DurationPredictor
    torch.manual_seed(cond + 3)
    m = DurationPredictor(192, 256, 3, 0.5, cond_channels=cond).eval()
    for p in m.parameters():
        p.data.add_(torch.randn_like(p) * 0.05)
    x = torch.randn(3, 192, 57)
    mask = (torch.arange(57)[None, :] < torch.tensor([57, 30, 1])[:, None]).float().unsqueeze(1)
    g = torch.randn(3, cond, 1) if cond else None
    want = O.duration_predictor(m.state_dict(), x, mask, g=g)
    m.cuda()
    got = m(x.cuda(), mask.cuda(), g=None if g is None else g.cuda())
    _close(got, want, 5e-5, "logw")


@pytest.mark.parametrize("cond", [0, 256])
def test_flow_full_width_vs_oracle(cond):
    from tts_b200.layers import ResidualCouplingBlocks
    torch.manual_seed(cond + 1)
    m = ResidualCouplingBlocks(192, 192, 5, 1, 4, cond_channels=cond).eval()
    _perturb(m)
    z = torch.randn(3, 192, 150)
    mask = O.sequence_mask(torch.tensor([150, 77, 3]), 150).unsqueeze(1).float()
    g = torch.randn(3, cond, 1) if cond else None
    want = O.flow_forward(m.state_dict(), z, mask, g, reverse=True)
    m.cuda()
    got = m(z.cuda(), mask.cuda(), g=None if g is None else g.cuda(), reverse=True)
    _close(got, want, 5e-5, "flow")
    # the flow is a bijection: forward(oracle) of the CUDA reverse returns the input on the valid region
    back = O.flow_forward(m.cpu().state_dict(), got.cpu(), mask, g, reverse=False)
    _close(back * mask, z * mask, 2e-4, "round trip")


def test_sdp_golden(golden):
    from tts_b200.layers import StochasticDurationPredictor
    g = golden("sdp_small")
    m = StochasticDurationPredictor(**g["args"]).eval()
    m.load_state_dict(g["state"])
    m.cuda()
    logw = m(g["x"].cuda(), g["x_mask"].cuda(), g=g["g"].cuda(), reverse=True, noise_scale=g["noise_scale"],
             noise=g["noise"])
    _close(logw, g["logw"], 5e-5, "logw")


@pytest.mark.parametrize("cond", [0, 256])
def test_sdp_full_width_vs_oracle(cond):
    from tts_b200.layers import StochasticDurationPredictor
    torch.manual_seed(3 + cond)
    m = StochasticDurationPredictor(192, 192, 3, 0.5, 4, cond_channels=cond).eval()
    _perturb(m)
    b, t = 4, 64
    x = torch.randn(b, 192, t)
    lens = torch.tensor([64, 40, 7, 1])
    mask = O.sequence_mask(lens, t).unsqueeze(1).float()
    x = x * mask
    noise = torch.randn(b, 2, t) * 2.5  # wide enough to land some samples in the linear tails (|x| > 5)
    g = torch.randn(b, cond, 1) if cond else None
    want = O.sdp_reverse(m.state_dict(), x, mask, noise, g=g, noise_scale=1.0)
    m.cuda()
    got = m(x.cuda(), mask.cuda(), g=None if g is None else g.cuda(), reverse=True, noise_scale=1.0, noise=noise)
    _close(got, want, 1e-4, "logw")
    assert int(m.last_error_flag.item()) == 0


def test_durations_and_path_expansion_bit_exact_vs_oracle():
    """Given the same logw, durations / lengths / alignment indices are bit-exact and the expanded prior
    (a gather) is exact; z_p differs only by the exp() implementation."""
    from tts_b200.layers import durations_to_path, expand_prior
    torch.manual_seed(5)
    b, tx, c = 6, 37, 24
    lens = torch.tensor([37, 30, 12, 5, 1, 37])
    mask = O.sequence_mask(lens, tx).unsqueeze(1).float()
    logw = torch.randn(b, 1, tx) * 0.8
    logw[5] = -20.0  # all durations round up to 1 only through ceil of tiny values
    stats = torch.randn(b, 2 * c, tx) * mask
    for length_scale in (1.0, 2.7):
        w = torch.exp(logw) * mask * length_scale
        w_ceil = torch.ceil(w)
        y_len = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        y_mask = O.sequence_mask(y_len, None).unsqueeze(1).float()
        attn = O.generate_path(w_ceil.squeeze(1), (mask * y_mask.transpose(1, 2)).squeeze(1).transpose(1, 2))
        m_p = torch.matmul(attn.transpose(1, 2), stats[:, :c].transpose(1, 2)).transpose(1, 2)
        logs_p = torch.matmul(attn.transpose(1, 2), stats[:, c:].transpose(1, 2)).transpose(1, 2)
        noise = torch.randn_like(m_p)
        z_p = m_p + noise * torch.exp(logs_p) * 0.667
        gw, gcum, gy, meta = durations_to_path(logw.cuda(), mask.cuda(), length_scale)
        assert torch.equal(gw.cpu(), w_ceil) and torch.equal(gy.cpu(), y_len)
        assert meta.tolist() == [int(y_len.max()), 0]
        t_dec = int(gy.max())
        ga, gm, gl, gz, gym = expand_prior(gcum, mask.cuda(), gy, stats.cuda(), noise.cuda(), 0.667, t_dec)
        assert torch.equal(ga.cpu(), attn) and torch.equal(gym.cpu(), y_mask)
        assert torch.equal(gm.cpu(), m_p) and torch.equal(gl.cpu(), logs_p)
        _close(gz, z_p, 1e-5, "z_p")
