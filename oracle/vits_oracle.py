"""TEST INFRASTRUCTURE ONLY -- CPU (PyTorch fp32) restatement of the coqui-ai/TTS
VITS + HiFiGAN inference hot path, written functionally over a reference-format
``state_dict``.  It is the checker for the CUDA path, never the thing shipped:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this file.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The port is pinned against the *unmodified* reference modules
(imported through oracle/ref_import.py in the build container) by
tests/test_oracle_vs_reference.py and against the committed fixtures under
tests/golden/ (generated from the real reference by tests/golden/make_golden.py).

Third-party arithmetic absent from /root/reference: ``librosa.filters.mel``
(librosa>=0.10, requirements.txt:9).  ``slaney_mel_basis`` restates its published
algorithm; the exact librosa values are **parity unpinned** (no reference test or
fixture pins them; cross-checked against torchaudio.functional.melscale_fbanks).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # vocoder/models/hifigan_generator.py:11


# --------------------------------------------------------------------------- weights
def conv_weight(sd, name):
    """Weight of a (possibly weight-normed) conv: w = g * v / ||v|| over dims != 0,
    the default ``dim=0`` of torch.nn.utils.parametrizations.weight_norm
    (hifigan_generator.py:36-80, wavenet.py:66-91).  Also accepts the legacy
    ``weight_g/weight_v`` keys of old checkpoints."""
    if name + ".weight" in sd:
        return sd[name + ".weight"]
    if name + ".parametrizations.weight.original0" in sd:
        g = sd[name + ".parametrizations.weight.original0"]
        v = sd[name + ".parametrizations.weight.original1"]
    else:
        g = sd[name + ".weight_g"]
        v = sd[name + ".weight_v"]
    return torch._weight_norm(v, g, 0)


def conv_bias(sd, name):
    return sd.get(name + ".bias", None)


def sub(sd, prefix):
    """View of a state dict below ``prefix.``"""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


# --------------------------------------------------------------------------- HiFiGAN
def hifigan_forward(sd, x, g=None, *, upsample_factors=(8, 8, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4),
                    resblock_kernel_sizes=(3, 7, 11), resblock_dilation_sizes=((1, 3, 5),) * 3,
                    resblock_type="1"):
    """HifiganGenerator.forward, vocoder/models/hifigan_generator.py:236-265;
    ResBlock1.forward :84-99, ResBlock2.forward :150-155."""
    o = F.conv1d(x, conv_weight(sd, "conv_pre"), conv_bias(sd, "conv_pre"), padding=3)
    if "cond_layer.weight" in sd:
        o = o + F.conv1d(g, sd["cond_layer.weight"], sd["cond_layer.bias"])
    nk = len(resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(upsample_factors, upsample_kernel_sizes)):
        o = F.leaky_relu(o, LRELU_SLOPE)
        o = F.conv_transpose1d(o, conv_weight(sd, f"ups.{i}"), conv_bias(sd, f"ups.{i}"), stride=u,
                               padding=(k - u) // 2)
        acc = None
        for j, (rk, rd) in enumerate(zip(resblock_kernel_sizes, resblock_dilation_sizes)):
            r = _resblock(sub(sd, f"resblocks.{i * nk + j}"), o, rk, rd, resblock_type)
            acc = r if acc is None else acc + r
        o = acc / nk
    o = F.leaky_relu(o)  # default slope 0.01 -- hifigan_generator.py:262
    o = F.conv1d(o, conv_weight(sd, "conv_post"), conv_bias(sd, "conv_post"), padding=3)
    return torch.tanh(o)


def _resblock(sd, x, k, dils, rtype):
    if rtype == "1":
        for n, d in enumerate(dils):
            t = F.leaky_relu(x, LRELU_SLOPE)
            t = F.conv1d(t, conv_weight(sd, f"convs1.{n}"), conv_bias(sd, f"convs1.{n}"), dilation=d,
                         padding=(k * d - d) // 2)
            t = F.leaky_relu(t, LRELU_SLOPE)
            t = F.conv1d(t, conv_weight(sd, f"convs2.{n}"), conv_bias(sd, f"convs2.{n}"), padding=(k - 1) // 2)
            x = t + x
        return x
    for n, d in enumerate(dils):
        t = F.leaky_relu(x, LRELU_SLOPE)
        t = F.conv1d(t, conv_weight(sd, f"convs.{n}"), conv_bias(sd, f"convs.{n}"), dilation=d,
                     padding=(k * d - d) // 2)
        x = t + x
    return x


def hifigan_inference(sd, c, inference_padding=5, **kw):
    """HifiganGenerator.inference :267-282 (replicate padding then forward)."""
    c = F.pad(c, (inference_padding, inference_padding), "replicate")
    return hifigan_forward(sd, c, **kw)


# --------------------------------------------------------------------------- WN + flow
def wn_forward(sd, x, x_mask, g=None, *, hidden, kernel_size, dilation_rate, num_layers):
    """WN.forward, tts/layers/generic/wavenet.py:94-115 (+ fused gate :6-13)."""
    out = torch.zeros_like(x)
    if g is not None:
        g = F.conv1d(g, conv_weight(sd, "cond_layer"), conv_bias(sd, "cond_layer"))
    for i in range(num_layers):
        d = dilation_rate ** i
        a = F.conv1d(x, conv_weight(sd, f"in_layers.{i}"), conv_bias(sd, f"in_layers.{i}"), dilation=d,
                     padding=(kernel_size * d - d) // 2)
        if g is not None:
            a = a + g[:, i * 2 * hidden:(i + 1) * 2 * hidden]
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])
        rs = F.conv1d(acts, conv_weight(sd, f"res_skip_layers.{i}"), conv_bias(sd, f"res_skip_layers.{i}"))
        if i < num_layers - 1:
            x = (x + rs[:, :hidden]) * x_mask
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out * x_mask


def coupling_block(sd, x, x_mask, g, reverse, *, hidden, kernel_size, dilation_rate, num_layers):
    """ResidualCouplingBlock.forward (mean_only=True), tts/layers/vits/networks.py:138-166."""
    half = x.shape[1] // 2
    x0, x1 = x[:, :half], x[:, half:]
    h = F.conv1d(x0, sd["pre.weight"], sd["pre.bias"]) * x_mask
    h = wn_forward(sub(sd, "enc"), h, x_mask, g, hidden=hidden, kernel_size=kernel_size,
                   dilation_rate=dilation_rate, num_layers=num_layers)
    m = F.conv1d(h, sd["post.weight"], sd["post.bias"]) * x_mask
    if reverse:
        x1 = (x1 - m) * x_mask
    else:
        x1 = m + x1 * x_mask
    return torch.cat([x0, x1], 1)


def flow_forward(sd, x, x_mask, g=None, reverse=False, *, num_flows=4, hidden=192, kernel_size=5,
                 dilation_rate=1, num_layers=4):
    """ResidualCouplingBlocks.forward, networks.py:214-232."""
    kw = dict(hidden=hidden, kernel_size=kernel_size, dilation_rate=dilation_rate, num_layers=num_layers)
    if not reverse:
        for n in range(num_flows):
            x = coupling_block(sub(sd, f"flows.{n}"), x, x_mask, g, False, **kw)
            x = torch.flip(x, [1])
    else:
        for n in reversed(range(num_flows)):
            x = torch.flip(x, [1])
            x = coupling_block(sub(sd, f"flows.{n}"), x, x_mask, g, True, **kw)
    return x


def posterior_encoder(sd, y, y_lengths, g=None, noise=None, *, out_channels=192, hidden=192, kernel_size=5,
                      dilation_rate=1, num_layers=16):
    """PosteriorEncoder.forward, networks.py:275-288 (noise supplied by the caller)."""
    y_mask = sequence_mask(y_lengths, y.shape[2]).unsqueeze(1).to(y.dtype)
    h = F.conv1d(y, sd["pre.weight"], sd["pre.bias"]) * y_mask
    h = wn_forward(sub(sd, "enc"), h, y_mask, g, hidden=hidden, kernel_size=kernel_size,
                   dilation_rate=dilation_rate, num_layers=num_layers)
    stats = F.conv1d(h, sd["proj.weight"], sd["proj.bias"]) * y_mask
    mean, log_scale = stats[:, :out_channels], stats[:, out_channels:]
    if noise is None:
        noise = torch.randn_like(mean)
    z = (mean + noise * torch.exp(log_scale)) * y_mask
    return z, mean, log_scale, y_mask


# --------------------------------------------------------------------------- text encoder
def sequence_mask(lengths, max_len=None):
    """tts/utils/helpers.py:43-57."""
    if max_len is None:
        max_len = lengths.max()          # a 0-dim tensor, like the reference: float lengths give ceil(max) columns
    return torch.arange(max_len, dtype=lengths.dtype, device=lengths.device)[None, :] < lengths[:, None]


def rel_attention(sd, x, attn_mask, *, num_heads, window):
    """RelativePositionMultiHeadAttention.forward/attention in closed form
    (tts/layers/glow_tts/transformer.py:109-163,196-241):
    scores_ij = (q_i.k_j + [|j-i|<=w] q_i.Ek[j-i+w]) / sqrt(d);  masked_fill(-1e4);
    softmax;  out_i = sum_j p_ij v_j + sum_{|j-i|<=w} p_ij Ev[j-i+w]."""
    b, c, t = x.shape
    d = c // num_heads
    q = F.conv1d(x, sd["conv_q.weight"], sd["conv_q.bias"]).view(b, num_heads, d, t).transpose(2, 3)
    k = F.conv1d(x, sd["conv_k.weight"], sd["conv_k.bias"]).view(b, num_heads, d, t).transpose(2, 3)
    v = F.conv1d(x, sd["conv_v.weight"], sd["conv_v.bias"]).view(b, num_heads, d, t).transpose(2, 3)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d)
    if window is not None:
        ek, ev = sd["emb_rel_k"], sd["emb_rel_v"]  # [1 or H, 2w+1, d]
        rel = torch.matmul(q, ek.unsqueeze(0).transpose(-2, -1)) / math.sqrt(d)  # [b,h,t,2w+1]
        idx = torch.arange(t, device=x.device)
        off = idx[None, :] - idx[:, None] + window  # j - i + w
        valid = (off >= 0) & (off <= 2 * window)
        gathered = torch.gather(rel, 3, off.clamp(0, 2 * window).expand(b, num_heads, t, t))
        scores = scores + torch.where(valid, gathered, torch.zeros_like(gathered))
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    p = F.softmax(scores, dim=-1)
    out = torch.matmul(p, v)
    if window is not None:
        pw = torch.zeros(b, num_heads, t, 2 * window + 1, dtype=p.dtype, device=p.device)
        for r in range(2 * window + 1):
            j = idx + (r - window)
            ok = (j >= 0) & (j < t)
            pw[:, :, ok, r] = p[:, :, idx[ok], j[ok]]
        out = out + torch.matmul(pw, ev.unsqueeze(0))
    out = out.transpose(2, 3).contiguous().view(b, c, t)
    return F.conv1d(out, sd["conv_o.weight"], sd["conv_o.bias"])


def layer_norm2(sd, x, eps=1e-5):
    """LayerNorm2, tts/layers/generic/normalization.py:31-53."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), sd["gamma"], sd["beta"], eps).transpose(1, -1)


def ffn(sd, x, x_mask, k):
    """FeedForwardNetwork.forward with _same_padding, transformer.py:290-313."""
    pl, pr = (k - 1) // 2, k // 2
    h = F.conv1d(F.pad(x * x_mask, (pl, pr)), sd["conv_1.weight"], sd["conv_1.bias"])
    h = torch.relu(h)
    h = F.conv1d(F.pad(h * x_mask, (pl, pr)), sd["conv_2.weight"], sd["conv_2.bias"])
    return h * x_mask


def rel_transformer(sd, x, x_mask, *, num_layers, num_heads, kernel_size, window=4):
    """RelativePositionTransformer.forward, transformer.py:411-432 (layer_norm_type "2")."""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    for i in range(num_layers):
        x = x * x_mask
        y = rel_attention(sub(sd, f"attn_layers.{i}"), x, attn_mask, num_heads=num_heads, window=window)
        x = layer_norm2(sub(sd, f"norm_layers_1.{i}"), x + y)
        y = ffn(sub(sd, f"ffn_layers.{i}"), x, x_mask, kernel_size)
        if i + 1 == num_layers and "proj.weight" in sd:
            x = F.conv1d(x, sd["proj.weight"], sd["proj.bias"])
        x = layer_norm2(sub(sd, f"norm_layers_2.{i}"), x + y)
    return x * x_mask


def text_encoder(sd, tokens, x_lengths, lang_emb=None, *, hidden=192, out_channels=192, num_heads=2,
                 num_layers=6, kernel_size=3):
    """TextEncoder.forward, tts/layers/vits/networks.py:80-100."""
    assert tokens.shape[0] == x_lengths.shape[0]
    x = F.embedding(tokens, sd["emb.weight"]) * math.sqrt(hidden)
    if lang_emb is not None:
        x = torch.cat((x, lang_emb.transpose(2, 1).expand(x.size(0), x.size(1), -1)), dim=-1)
    x = x.transpose(1, -1)
    x_mask = sequence_mask(x_lengths, x.size(2)).unsqueeze(1).to(x.dtype)
    x = rel_transformer(sub(sd, "encoder"), x * x_mask, x_mask, num_layers=num_layers, num_heads=num_heads,
                        kernel_size=kernel_size)
    stats = F.conv1d(x, sd["proj.weight"], sd["proj.bias"]) * x_mask
    return x, stats[:, :out_channels], stats[:, out_channels:], x_mask


# --------------------------------------------------------------------------- SDP
def dds_conv(sd, x, x_mask, g=None, *, kernel_size=3, num_layers=3):
    """DilatedDepthSeparableConv.forward, stochastic_duration_predictor.py:46-63."""
    if g is not None:
        x = x + g
    c = x.shape[1]
    for i in range(num_layers):
        d = kernel_size ** i
        y = F.conv1d(x * x_mask, sd[f"convs_sep.{i}.weight"], sd[f"convs_sep.{i}.bias"], groups=c, dilation=d,
                     padding=(kernel_size * d - d) // 2)
        y = F.gelu(layer_norm2(sub(sd, f"norms_1.{i}"), y))
        y = F.conv1d(y, sd[f"convs_1x1.{i}.weight"], sd[f"convs_1x1.{i}.bias"])
        y = F.gelu(layer_norm2(sub(sd, f"norms_2.{i}"), y))
        x = x + y
    return x * x_mask


def rq_spline_inverse(x, uw, uh, ud, tail_bound=5.0, min_w=1e-3, min_h=1e-3, min_d=1e-3):
    """Inverse of the unconstrained rational-quadratic spline with linear tails,
    tts/layers/vits/transforms.py:51-184 (inverse branch :142,159-171)."""
    nb = uw.shape[-1]
    inside = (x >= -tail_bound) & (x <= tail_bound)
    const = float(np.log(np.exp(1 - min_d) - 1))
    ud = F.pad(ud, (1, 1))
    ud[..., 0] = const
    ud[..., -1] = const
    out = x.clone()
    if not inside.any():
        return out
    xi, uw, uh, ud = x[inside], uw[inside], uh[inside], ud[inside]
    left = bottom = -tail_bound
    right = top = tail_bound
    w = min_w + (1 - min_w * nb) * F.softmax(uw, dim=-1)
    cw = F.pad(torch.cumsum(w, dim=-1), (1, 0))
    cw = (right - left) * cw + left
    cw[..., 0], cw[..., -1] = left, right
    w = cw[..., 1:] - cw[..., :-1]
    dv = min_d + F.softplus(ud)
    h = min_h + (1 - min_h * nb) * F.softmax(uh, dim=-1)
    ch = F.pad(torch.cumsum(h, dim=-1), (1, 0))
    ch = (top - bottom) * ch + bottom
    ch[..., 0], ch[..., -1] = bottom, top
    h = ch[..., 1:] - ch[..., :-1]
    loc = ch.clone()
    loc[..., -1] += 1e-6  # searchsorted eps, transforms.py:45-47
    b = (torch.sum(xi[..., None] >= loc, dim=-1) - 1)[..., None]
    in_cw, in_w = cw.gather(-1, b)[..., 0], w.gather(-1, b)[..., 0]
    in_ch, in_h = ch.gather(-1, b)[..., 0], h.gather(-1, b)[..., 0]
    in_delta = (h / w).gather(-1, b)[..., 0]
    d0, d1 = dv.gather(-1, b)[..., 0], dv[..., 1:].gather(-1, b)[..., 0]
    qa = (xi - in_ch) * (d0 + d1 - 2 * in_delta) + in_h * (in_delta - d0)
    qb = in_h * d0 - (xi - in_ch) * (d0 + d1 - 2 * in_delta)
    qc = -in_delta * (xi - in_ch)
    disc = qb.pow(2) - 4 * qa * qc
    assert (disc >= 0).all()
    root = (2 * qc) / (-qb - torch.sqrt(disc))
    out[inside] = root * in_w + in_cw
    return out


def conv_flow_reverse(sd, z, x_mask, g, *, hidden=192, num_bins=10, tail_bound=5.0, kernel_size=3):
    """ConvFlow.forward(reverse=True), stochastic_duration_predictor.py:120-147."""
    x0, x1 = z[:, :1], z[:, 1:]
    h = F.conv1d(x0, sd["pre.weight"], sd["pre.bias"])
    h = dds_conv(sub(sd, "convs"), h, x_mask, g=g, kernel_size=kernel_size, num_layers=3)
    h = F.conv1d(h, sd["proj.weight"], sd["proj.bias"]) * x_mask
    b, c, t = x0.shape
    h = h.reshape(b, c, -1, t).permute(0, 1, 3, 2)
    uw = h[..., :num_bins] / math.sqrt(hidden)
    uh = h[..., num_bins:2 * num_bins] / math.sqrt(hidden)
    ud = h[..., 2 * num_bins:]
    x1 = rq_spline_inverse(x1, uw, uh, ud, tail_bound=tail_bound)
    return torch.cat([x0, x1], 1) * x_mask


def sdp_reverse(sd, x, x_mask, noise, g=None, lang_emb=None, noise_scale=1.0, *, hidden=192, kernel_size=3,
                num_flows=4):
    """StochasticDurationPredictor.forward(reverse=True), stochastic_duration_predictor.py:222-239,285-294.
    ``noise`` is the [B,2,T] standard-normal draw of :287 (made on the CPU generator by the caller)."""
    x = F.conv1d(x, sd["pre.weight"], sd["pre.bias"])
    if g is not None:
        x = x + F.conv1d(g, sd["cond.weight"], sd["cond.bias"])
    if lang_emb is not None:
        x = x + F.conv1d(lang_emb, sd["cond_lang.weight"], sd["cond_lang.bias"])
    x = dds_conv(sub(sd, "convs"), x, x_mask, kernel_size=kernel_size, num_layers=3)
    x = F.conv1d(x, sd["proj.weight"], sd["proj.bias"]) * x_mask
    order = list(reversed(range(num_flows + 1)))  # flows[0] is the ElementwiseAffine
    order = order[:-2] + [order[-1]]  # "remove a useless vflow" :286
    z = noise.to(x.dtype) * noise_scale
    for n in order:
        z = torch.flip(z, [1])
        f = sub(sd, f"flows.{n}")
        if n == 0:  # ElementwiseAffine reverse :83
            z = (z - f["translation"]) * torch.exp(-f["log_scale"]) * x_mask
        else:
            z = conv_flow_reverse(f, z, x_mask, x, hidden=hidden, kernel_size=kernel_size)
    return z[:, :1]


def duration_predictor(sd, x, x_mask, g=None, lang_emb=None):
    """Deterministic DurationPredictor.forward, tts/layers/glow_tts/duration_predictor.py:44-69
    (LayerNorm with eps 1e-4 over channels, normalization.py:5-28)."""
    def ln(p, v):
        mean = v.mean(1, keepdim=True)
        var = ((v - mean) ** 2).mean(1, keepdim=True)
        return (v - mean) * torch.rsqrt(var + 1e-4) * p["gamma"] + p["beta"]

    if g is not None:
        x = x + F.conv1d(g, sd["cond.weight"], sd["cond.bias"])
    if lang_emb is not None:
        x = x + F.conv1d(lang_emb, sd["cond_lang.weight"], sd["cond_lang.bias"])
    k = sd["conv_1.weight"].shape[-1]
    x = ln(sub(sd, "norm_1"), torch.relu(F.conv1d(x * x_mask, sd["conv_1.weight"], sd["conv_1.bias"], padding=k // 2)))
    x = ln(sub(sd, "norm_2"), torch.relu(F.conv1d(x * x_mask, sd["conv_2.weight"], sd["conv_2.bias"], padding=k // 2)))
    return F.conv1d(x * x_mask, sd["proj.weight"], sd["proj.bias"]) * x_mask


# --------------------------------------------------------------------------- durations -> path
def generate_path(duration, mask):
    """tts/utils/helpers.py:154-169.  duration [B,Tx]; mask [B,Tx,Ty]."""
    b, t_x, t_y = mask.shape
    cum = torch.cumsum(duration, 1).view(b * t_x)
    path = sequence_mask(cum, t_y).to(mask.dtype).view(b, t_x, t_y)
    path = path - F.pad(path, (0, 0, 1, 0))[:, :-1]
    return path * mask


def vits_inference(sd, tokens, x_lengths, sdp_noise, prior_noise_fn, *, args, speaker_ids=None, d_vectors=None,
                   language_ids=None):
    """Vits.inference glue, tts/models/vits.py:1112-1173 (+ _set_cond_input :874-894).

    ``sdp_noise``: [B,2,Tt] CPU standard normal (stochastic_duration_predictor.py:287).
    ``prior_noise_fn(shape)``: returns the randn_like(m_p) draw of vits.py:1155.
    ``args``: dict of VitsArgs fields (vits.py:544-600).
    """
    a = args
    g = None
    if a.get("use_speaker_embedding") and speaker_ids is not None:
        g = F.embedding(speaker_ids, sd["emb_g.weight"]).unsqueeze(-1)
    elif d_vectors is not None:
        g = F.normalize(d_vectors).unsqueeze(-1)
    lang_emb = None
    if a.get("use_language_embedding") and language_ids is not None:
        lang_emb = F.embedding(language_ids, sd["emb_l.weight"]).unsqueeze(-1)
    hid = a["hidden_channels"]
    x, m_p, logs_p, x_mask = text_encoder(
        sub(sd, "text_encoder"), tokens, x_lengths, lang_emb, hidden=hid, out_channels=hid,
        num_heads=a["num_heads_text_encoder"], num_layers=a["num_layers_text_encoder"],
        kernel_size=a["kernel_size_text_encoder"])
    g_dp = g if a.get("condition_dp_on_speaker", True) else None
    if a.get("use_sdp", True):
        logw = sdp_reverse(sub(sd, "duration_predictor"), x, x_mask, sdp_noise, g=g_dp, lang_emb=lang_emb,
                           noise_scale=a.get("inference_noise_scale_dp", 1.0))
    else:
        logw = duration_predictor(sub(sd, "duration_predictor"), x, x_mask, g=g_dp, lang_emb=lang_emb)
    w = torch.exp(logw) * x_mask * a.get("length_scale", 1.0)
    w_ceil = torch.ceil(w)
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    y_mask = sequence_mask(y_lengths, None).to(x_mask.dtype).unsqueeze(1)
    attn_mask = x_mask * y_mask.transpose(1, 2)
    attn = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1).transpose(1, 2))
    m_p = torch.matmul(attn.transpose(1, 2), m_p.transpose(1, 2)).transpose(1, 2)
    logs_p = torch.matmul(attn.transpose(1, 2), logs_p.transpose(1, 2)).transpose(1, 2)
    noise = prior_noise_fn(m_p.shape)
    z_p = m_p + noise * torch.exp(logs_p) * a.get("inference_noise_scale", 0.667)
    z = flow_forward(sub(sd, "flow"), z_p, y_mask, g=g, reverse=True, hidden=hid,
                     kernel_size=a["kernel_size_flow"], dilation_rate=a["dilation_rate_flow"],
                     num_layers=a["num_layers_flow"])
    if a.get("encoder_sample_rate") and a.get("interpolate_z", True):   # upsampling_z, vits.py:944-959
        f = a["sample_rate"] / a["encoder_sample_rate"]
        z = F.interpolate(z, scale_factor=[f], mode="linear").squeeze(0)
        y_mask = sequence_mask(y_lengths * f, None).to(y_mask.dtype).unsqueeze(1)
    mil = a.get("max_inference_len", None)
    o = hifigan_forward(sub(sd, "waveform_decoder"), (z * y_mask)[:, :, :mil], g=g,
                        upsample_factors=a["upsample_rates_decoder"],
                        upsample_kernel_sizes=a["upsample_kernel_sizes_decoder"],
                        resblock_kernel_sizes=a["resblock_kernel_sizes_decoder"],
                        resblock_dilation_sizes=a["resblock_dilation_sizes_decoder"],
                        resblock_type=a["resblock_type_decoder"])
    return {"model_outputs": o, "alignments": attn, "durations": w_ceil, "z": z, "z_p": z_p, "m_p": m_p,
            "logs_p": logs_p, "y_mask": y_mask, "logw": logw, "x": x, "y_lengths": y_lengths}


def voice_conversion(sd, y, y_lengths, g_src, g_tgt, posterior_noise, *, args):
    """Vits.voice_conversion glue, tts/models/vits.py:1226-1232 (g_* already embedded / normalised [B,C,1])."""
    a = args
    hid = a["hidden_channels"]
    z, _, _, y_mask = posterior_encoder(sub(sd, "posterior_encoder"), y, y_lengths, g=g_src, noise=posterior_noise,
                                        out_channels=hid, hidden=hid,
                                        kernel_size=a["kernel_size_posterior_encoder"],
                                        dilation_rate=a["dilation_rate_posterior_encoder"],
                                        num_layers=a["num_layers_posterior_encoder"])
    fkw = dict(hidden=hid, kernel_size=a["kernel_size_flow"], dilation_rate=a["dilation_rate_flow"],
               num_layers=a["num_layers_flow"])
    z_p = flow_forward(sub(sd, "flow"), z, y_mask, g=g_src, reverse=False, **fkw)
    z_hat = flow_forward(sub(sd, "flow"), z_p, y_mask, g=g_tgt, reverse=True, **fkw)
    o_hat = hifigan_forward(sub(sd, "waveform_decoder"), z_hat * y_mask, g=g_tgt,
                            upsample_factors=a["upsample_rates_decoder"],
                            upsample_kernel_sizes=a["upsample_kernel_sizes_decoder"],
                            resblock_kernel_sizes=a["resblock_kernel_sizes_decoder"],
                            resblock_dilation_sizes=a["resblock_dilation_sizes_decoder"],
                            resblock_type=a["resblock_type_decoder"])
    return o_hat, y_mask, (z, z_p, z_hat)


# --------------------------------------------------------------------------- MAS
def maximum_path_numpy_loop(value, t_xs, t_ys, max_neg_val=-1e9):
    """Pure-Python/numpy transcription of core.pyx:11-37 for SMALL cases only."""
    value = np.array(value, dtype=np.float32, copy=True)
    b, tx, ty = value.shape
    path = np.zeros((b, tx, ty), dtype=np.int32)
    neg = np.float32(max_neg_val)
    for n in range(b):
        t_x, t_y = int(t_xs[n]), int(t_ys[n])
        v = value[n]
        for y in range(t_y):
            for x in range(max(0, t_x + y - t_y), min(t_x, y + 1)):
                v_cur = neg if x == y else v[x, y - 1]
                v_prev = (np.float32(0.0) if y == 0 else neg) if x == 0 else v[x - 1, y - 1]
                v[x, y] = np.float32(max(v_cur, v_prev) + v[x, y])
        index = t_x - 1
        for y in range(t_y - 1, -1, -1):
            path[n, index, y] = 1
            if index != 0 and y > 0 and (index == y or v[index, y - 1] < v[index - 1, y - 1]):
                index -= 1
    return path


_MAS_LIB = None


def _mas_lib():
    global _MAS_LIB
    if _MAS_LIB is None:
        import ctypes
        import os
        here = os.path.dirname(os.path.abspath(__file__))
        _MAS_LIB = ctypes.CDLL(os.path.join(here, "_build", "libmas_oracle.so"))
    return _MAS_LIB


def maximum_path_c_port(value, t_xs, t_ys, max_neg_val=-1e9):
    """oracle/mas_oracle.c (the C restatement of core.pyx) through ctypes."""
    import ctypes
    v = np.ascontiguousarray(np.array(value, dtype=np.float32, copy=True))
    b, tx, ty = v.shape
    path = np.zeros((b, tx, ty), dtype=np.int32)
    txs = np.ascontiguousarray(t_xs, dtype=np.int32)
    tys = np.ascontiguousarray(t_ys, dtype=np.int32)
    _mas_lib().mas_oracle_f32(path.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p),
                              txs.ctypes.data_as(ctypes.c_void_p), tys.ctypes.data_as(ctypes.c_void_p),
                              ctypes.c_int(b), ctypes.c_int(tx), ctypes.c_int(ty), ctypes.c_float(max_neg_val))
    return path


def maximum_path(value, mask, impl="c"):
    """maximum_path / maximum_path_cython, tts/utils/helpers.py:172-194 (torch in, torch out)."""
    value = value * mask
    dtype = value.dtype
    v = value.detach().cpu().numpy().astype(np.float32)
    m = mask.detach().cpu().numpy()
    t_x = m.sum(1)[:, 0].astype(np.int32)
    t_y = m.sum(2)[:, 0].astype(np.int32)
    if impl == "ref":
        from ref_import import load_ref_mas_core  # oracle/_ref compiled reference kernel
        core = load_ref_mas_core()
        path = np.zeros_like(v).astype(np.int32)
        core.maximum_path_c(path, v, t_x, t_y)
    elif impl == "py":
        path = maximum_path_numpy_loop(v, t_x, t_y)
    else:
        path = maximum_path_c_port(v, t_x, t_y)
    return torch.from_numpy(path).to(dtype=dtype)


# --------------------------------------------------------------------------- STFT / mel front end
def slaney_mel_basis(sample_rate, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney') restated from the
    published algorithm (used at tts/models/vits.py:153,180 and utils/audio/numpy_transforms.py:31).
    Returns float32 [n_mels, n_fft//2+1]."""
    if fmax is None:
        fmax = sample_rate / 2.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f_sp = 200.0 / 3
        freqs = f_sp * m
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)

    n_freq = 1 + n_fft // 2
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sample_rate)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_freq), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def wav_to_spec(y, n_fft, hop_length, win_length, center=False):
    """tts/models/vits.py:96-138: reflect-pad (n_fft-hop)/2, hann STFT, sqrt(re^2+im^2+1e-6)."""
    y = y.squeeze(1)
    pad = int((n_fft - hop_length) / 2)
    y = F.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop_length, win_length=win_length,
                      window=torch.hann_window(win_length, dtype=y.dtype), center=center, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    spec = torch.view_as_real(spec)
    return torch.sqrt(spec.pow(2).sum(-1) + 1e-6)


def spec_to_mel(spec, n_fft, num_mels, sample_rate, fmin, fmax):
    """tts/models/vits.py:141-157: mel basis @ spec, log(clamp(., 1e-5))."""
    basis = torch.from_numpy(slaney_mel_basis(sample_rate, n_fft, num_mels, fmin, fmax)).to(spec.dtype)
    return torch.log(torch.clamp(torch.matmul(basis, spec), min=1e-5))


def wav_to_mel(y, n_fft, num_mels, sample_rate, hop_length, win_length, fmin, fmax, center=False):
    """tts/models/vits.py:160-208 == spec_to_mel(wav_to_spec(.)) (tests/tts_tests/test_vits.py:56)."""
    return spec_to_mel(wav_to_spec(y, n_fft, hop_length, win_length, center), n_fft, num_mels, sample_rate, fmin,
                       fmax)


def torch_stft_call(x, n_fft, hop_length, win_length, *, pad_wav=False, power=None, use_mel=False,
                    mel_basis=None, do_amp_to_db=False, spec_gain=1.0, normalized=False):
    """TorchSTFT.__call__, utils/audio/torch_transforms.py:104-145 (center=True, clamp 1e-8)."""
    if x.ndim == 2:
        x = x.unsqueeze(1)
    if pad_wav:
        pad = int((n_fft - hop_length) / 2)
        x = F.pad(x, (pad, pad), mode="reflect")
    o = torch.stft(x.squeeze(1), n_fft, hop_length, win_length, torch.hann_window(win_length), center=True,
                   pad_mode="reflect", normalized=normalized, onesided=True, return_complex=True)
    o = torch.view_as_real(o)
    s = torch.sqrt(torch.clamp(o[..., 0] ** 2 + o[..., 1] ** 2, min=1e-8))
    if power is not None:
        s = s ** power
    if use_mel:
        s = torch.matmul(mel_basis.to(s), s)
    if do_amp_to_db:
        s = torch.log(torch.clamp(s, min=1e-5) * spec_gain)
    return s


# --------------------------------------------------------------------------- vocoder hand-off (Synthesizer.tts)
def audio_normalize(S, *, signal_norm=True, symmetric_norm=True, max_norm=4.0, clip_norm=True, min_level_db=-100.0,
                    ref_level_db=20.0, mel_mean=None, mel_std=None):
    """AudioProcessor.normalize, utils/audio/processor.py:259-301 (S float32 numpy [C,T]; the mean-var branch is
    StandardScaler.transform with the given per-channel statistics)."""
    S = np.array(S, dtype=np.float32, copy=True)
    if not signal_norm:
        return S
    if mel_mean is not None:
        return ((S.T - np.asarray(mel_mean, dtype=np.float32)) / np.asarray(mel_std, dtype=np.float32)).T
    S -= np.float32(ref_level_db)
    S_norm = (S - np.float32(min_level_db)) / np.float32(-min_level_db)
    if symmetric_norm:
        S_norm = (np.float32(2 * max_norm) * S_norm) - np.float32(max_norm)
        if clip_norm:
            S_norm = np.clip(S_norm, -max_norm, max_norm)
        return S_norm.astype(np.float32)
    S_norm = np.float32(max_norm) * S_norm
    if clip_norm:
        S_norm = np.clip(S_norm, 0, max_norm)
    return S_norm.astype(np.float32)


def audio_denormalize(S, *, signal_norm=True, symmetric_norm=True, max_norm=4.0, clip_norm=True, min_level_db=-100.0,
                      ref_level_db=20.0, mel_mean=None, mel_std=None):
    """AudioProcessor.denormalize, utils/audio/processor.py:303-337."""
    S = np.array(S, dtype=np.float32, copy=True)
    if not signal_norm:
        return S
    if mel_mean is not None:
        return (S.T * np.asarray(mel_std, dtype=np.float32) + np.asarray(mel_mean, dtype=np.float32)).T
    if symmetric_norm:
        if clip_norm:
            S = np.clip(S, -max_norm, max_norm)
        S = ((S + np.float32(max_norm)) * np.float32(-min_level_db) / np.float32(2 * max_norm)) + np.float32(min_level_db)
        return (S + np.float32(ref_level_db)).astype(np.float32)
    if clip_norm:
        S = np.clip(S, 0, max_norm)
    S = (S * np.float32(-min_level_db) / np.float32(max_norm)) + np.float32(min_level_db)
    return (S + np.float32(ref_level_db)).astype(np.float32)


def interpolate_vocoder_input(scale_factor, spec):
    """vocoder/utils/generic_utils.py:11-29: spec [C,T] -> [1,C,T']."""
    spec = torch.as_tensor(np.asarray(spec)).unsqueeze(0).unsqueeze(0)
    return F.interpolate(spec, scale_factor=scale_factor, recompute_scale_factor=True, mode="bilinear",
                         align_corners=False).squeeze(0)


def vocoder_handoff(mel_tc, tts_norm, vocoder_norm, sr_tts, sr_vocoder, inference_padding=5):
    """The chain of utils/synthesizer.py:412-429 for one sentence + the replicate pad of HifiganGenerator.inference
    (hifigan_generator.py:281): mel_tc is the TTS model output [T, C] (normalised by the TTS AudioProcessor);
    returns the tensor conv_pre sees, [1, C, T' + 2*pad]."""
    mel = audio_denormalize(np.asarray(mel_tc).T, **tts_norm).T          # .denormalize(mel.T).T
    vin = audio_normalize(mel.T, **vocoder_norm)                         # vocoder_ap.normalize(mel.T)
    scale = [1, sr_vocoder / sr_tts]
    if scale[1] != 1:
        vin = interpolate_vocoder_input(scale, vin)
    else:
        vin = torch.as_tensor(vin).unsqueeze(0)
    return F.pad(vin, (inference_padding, inference_padding), "replicate")


def wav_to_int16(wav):
    """save_wav's peak normalisation, utils/audio/numpy_transforms.py:439-441."""
    wav = np.asarray(wav, dtype=np.float32)
    wav_norm = wav * (32767 / max(0.01, np.max(np.abs(wav))))
    return wav_norm.astype(np.int16)


# --------------------------------------------------------------------------- training-side alignment
def forward_mas_attn(z_p, m_p, logs_p, x_mask, y_mask, impl="c"):
    """The alignment half of Vits.forward_mas, tts/models/vits.py:909-919: attn [B,1,Tx,Ty] (and logp)."""
    attn_mask = torch.unsqueeze(x_mask, -1) * torch.unsqueeze(y_mask, 2)
    o_scale = torch.exp(-2 * logs_p)
    logp1 = torch.sum(-0.5 * math.log(2 * math.pi) - logs_p, [1]).unsqueeze(-1)
    logp2 = torch.einsum("klm, kln -> kmn", [o_scale, -0.5 * (z_p ** 2)])
    logp3 = torch.einsum("klm, kln -> kmn", [m_p * o_scale, z_p])
    logp4 = torch.sum(-0.5 * (m_p ** 2) * o_scale, [1]).unsqueeze(-1)
    logp = logp2 + logp3 + logp1 + logp4
    attn = maximum_path(logp, attn_mask.squeeze(1), impl=impl).unsqueeze(1)
    return attn, logp
