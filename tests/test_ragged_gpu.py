"""Ragged batches (tts_b200 extension, ``Vits.trim_padding`` / ``lengths=``): padded frames are neither computed nor
read, and every sample below a row's length must be BIT-IDENTICAL to the dense computation -- checked with torch.equal
against the dense CUDA path (which the other tests tie to the oracle), plus zeros beyond the row's end."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _decoder(cond=0):
    from tts_b200.hifigan import HifiganGenerator
    return HifiganGenerator(in_channels=192, out_channels=1, resblock_type="1", resblock_dilation_sizes=[[1, 3, 5]] * 3,
                            resblock_kernel_sizes=[3, 7, 11], upsample_kernel_sizes=[16, 16, 4, 4],
                            upsample_initial_channel=512, upsample_factors=[8, 8, 2, 2], inference_padding=0,
                            cond_channels=cond, conv_pre_weight_norm=False, conv_post_weight_norm=False,
                            conv_post_bias=False).eval()


@pytest.mark.parametrize("b,t,lens", [(6, 150, [150, 97, 64, 33, 2, 1]), (3, 64, [64, 64, 10]), (2, 301, [17, 301]),
                                      (5, 40, [40, 0, 7, 40, 23])])
def test_decoder_ragged_equals_dense_on_valid_samples(b, t, lens):
    torch.manual_seed(b * 1000 + t)
    m = _decoder().cuda()
    lens_t = torch.tensor(lens)
    mask = (torch.arange(t)[None, :] < lens_t[:, None]).float().unsqueeze(1)
    z = (torch.randn(b, 192, t) * mask).cuda()                  # what Vits feeds: z * y_mask
    dense = m(z)
    torch.cuda.synchronize()
    ws_poison = torch.full((1 << 26,), float("nan"), device="cuda")   # recycled scratch must not leak into valid samples
    del ws_poison
    ragged = m(z, lengths=lens_t.cuda())
    assert ragged.shape == dense.shape
    for i, n in enumerate(lens):
        assert torch.equal(ragged[i, :, : n * 256], dense[i, :, : n * 256]), (i, n)
        assert float(ragged[i, :, n * 256:].abs().sum()) == 0.0, (i, n)     # the padded tail is clean zeros
    assert torch.isfinite(ragged).all()


def test_decoder_ragged_with_peak_and_conditioning():
    from tts_b200.vocoder import new_peak
    torch.manual_seed(3)
    m = _decoder(cond=256).cuda()
    lens = torch.tensor([90, 41, 12])
    mask = (torch.arange(90)[None, :] < lens[:, None]).float().unsqueeze(1)
    z, g = (torch.randn(3, 192, 90) * mask).cuda(), torch.randn(3, 256, 1).cuda()
    dense = m(z, g)
    peak = new_peak(z.device)
    ragged = m(z, g, peak=peak, lengths=lens.cuda())
    valid = (torch.arange(90 * 256)[None, None, :] < (lens * 256)[:, None, None]).cuda()
    assert torch.equal(ragged * valid, dense * valid)
    want_peak = (dense * valid).abs().max()
    assert torch.equal(peak.view(torch.float32)[0], want_peak)


def test_flow_ragged_equals_dense():
    from tts_b200.layers import ResidualCouplingBlocks
    torch.manual_seed(5)
    fl = ResidualCouplingBlocks(192, 192, 5, 1, 4, cond_channels=256).eval()
    for _, p in fl.named_parameters():
        if float(p.detach().abs().sum()) == 0.0:
            p.data.normal_(0, 0.05)
    fl.cuda()
    lens = torch.tensor([301, 210, 140, 9, 1])
    mask = (torch.arange(301)[None, :] < lens[:, None]).float().unsqueeze(1).cuda()
    z_p, g = torch.randn(5, 192, 301).cuda(), torch.randn(5, 256, 1).cuda()
    dense = fl(z_p, mask, g=g, reverse=True)
    ragged = fl(z_p, mask, g=g, reverse=True, lengths=lens.cuda())
    assert torch.equal(ragged, dense)          # dense is zero beyond each row's end (per-layer masking), ragged is masked


def test_vits_trim_padding_bit_identical_on_valid_samples():
    from tts_b200.vits import Vits, VitsArgs, VitsConfig
    torch.manual_seed(8)
    cfg = VitsConfig(model_args=VitsArgs(use_speaker_embedding=True, num_speakers=9))
    m = Vits(cfg).eval()
    gen = torch.Generator().manual_seed(1)
    for _, p in m.named_parameters():
        if float(p.detach().abs().sum()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=gen) * 0.05)
    m.cuda()
    lens = torch.tensor([64, 50, 33, 20, 11, 64, 5, 41])
    tok = (torch.randint(0, 100, (8, 64)) * (torch.arange(64)[None, :] < lens[:, None])).cuda()
    aux = {"x_lengths": lens.cuda(), "speaker_ids": torch.randint(0, 9, (8,)).cuda()}
    noise = torch.randn(8, 2, 64)
    store = {}

    def prior(shape):
        if "n" not in store:
            store["n"] = torch.randn(shape, generator=torch.Generator().manual_seed(2)).cuda()
        return store["n"]

    dense = m.inference(tok, aux, sdp_noise=noise, prior_noise=prior)
    m.trim_padding = True
    ragged = m.inference(tok, aux, sdp_noise=noise, prior_noise=prior)
    assert torch.equal(ragged["wav_lengths"], dense["wav_lengths"]) and torch.equal(ragged["alignments"], dense["alignments"])
    wl = dense["wav_lengths"].tolist()
    assert min(wl) < max(wl)                                       # the batch really is ragged
    for i, n in enumerate(wl):
        assert torch.equal(ragged["model_outputs"][i, :, :n], dense["model_outputs"][i, :, :n]), i
        assert float(ragged["model_outputs"][i, :, n:].abs().sum()) == 0.0
    assert torch.equal(ragged["z"], dense["z"] * dense["y_mask"])
