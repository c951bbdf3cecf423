"""Drop-in for the hot functions of TTS/tts/utils/helpers.py (sequence_mask :43-57,
generate_path :154-169, maximum_path :172-194) on top of the CUDA library."""
import ctypes

import torch

from . import _lib


def sequence_mask(sequence_length, max_len=None):
    """Same contract as TTS.tts.utils.helpers.sequence_mask (helpers.py:43-57)."""
    if max_len is None:
        max_len = int(sequence_length.max())
    seq_range = torch.arange(max_len, dtype=sequence_length.dtype, device=sequence_length.device)
    return seq_range.unsqueeze(0) < sequence_length.unsqueeze(1)


def maximum_path_lengths(value, t_x, t_y, mask=None, out_dtype=None):
    """Monotonic alignment search given explicit lengths (the C-ABI form).

    value [B,Tx,Ty] f32 CUDA; t_x,t_y int32 [B] CUDA; mask optional [B,Tx,Ty] f32 (value*mask is
    formed on load).  Returns path [B,Tx,Ty] int32 (default, the dtype of core.pyx) or float32."""
    _lib.require_cuda(value, "value")
    if value.dtype != torch.float32:
        value = value.float()
    value = value.contiguous()
    b, tx, ty = value.shape
    f32 = out_dtype == torch.float32
    path = torch.empty((b, tx, ty), dtype=torch.float32 if f32 else torch.int32, device=value.device)
    if mask is not None:
        mask = mask.to(torch.float32).contiguous()
    t_x = t_x.to(device=value.device, dtype=torch.int32).contiguous()
    t_y = t_y.to(device=value.device, dtype=torch.int32).contiguous()
    L = _lib.lib()
    with torch.cuda.device(value.device):
        nbytes = L.b200tts_mas_workspace_bytes(b, tx, ty)
        ws = _lib.workspace(value.device, nbytes, "mas")
        rc = L.b200tts_mas(_lib.ptr(value), _lib.ptr(mask), _lib.ptr(t_x), _lib.ptr(t_y), b, tx, ty, _lib.ptr(path),
                           1 if f32 else 0, _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_ptr(value.device))
    _lib.check(rc, "mas")
    return path


def maximum_path(value, mask):
    """Same contract as TTS.tts.utils.helpers.maximum_path (helpers.py:172-194): value, mask
    [B,T_en,T_de]; returns the 0/1 path in value's dtype on value's device -- without the
    reference's device->host->device round trip."""
    _lib.require_cuda(value, "value")
    dtype = value.dtype
    maskf = mask.to(torch.float32)
    t_x = maskf[:, :, 0].sum(1).to(torch.int32)   # mask.sum(1)[:, 0]
    t_y = maskf[:, 0, :].sum(1).to(torch.int32)   # mask.sum(2)[:, 0]
    path = maximum_path_lengths(value, t_x, t_y, mask=maskf, out_dtype=torch.float32)
    return path if dtype == torch.float32 else path.to(dtype)


def generate_path(duration, mask):
    """Same contract as TTS.tts.utils.helpers.generate_path (helpers.py:154-169) -- torch ops; the
    fused CUDA form used by Vits.inference lives in tts_b200.vits."""
    b, t_x, t_y = mask.shape
    cum = torch.cumsum(duration, 1).view(b * t_x)
    path = sequence_mask(cum, t_y).to(mask.dtype).view(b, t_x, t_y)
    path = path - torch.nn.functional.pad(path, (0, 0, 1, 0))[:, :-1]
    return path * mask


def maximum_path_from_stats(z_p, m_p, logs_p, x_mask, y_mask, return_logp=False):
    """The alignment step of ``Vits.forward_mas`` (TTS/tts/models/vits.py:909-919) on the device: builds the
    log-likelihood ``logp`` [B,Tx,Ty] of every (text position, frame) pair from the prior statistics and runs the
    monotonic alignment search on it.  z_p [B,C,Ty]; m_p, logs_p [B,C,Tx]; x_mask [B,1,Tx]; y_mask [B,1,Ty].
    Returns ``attn`` [B,1,Tx,Ty] (0/1, float32) like the reference (and ``logp`` when asked)."""
    _lib.require_cuda(z_p, "z_p")
    z_p, m_p, logs_p = (t.to(torch.float32).contiguous() for t in (z_p, m_p, logs_p))
    b, c, ty = z_p.shape
    tx = m_p.shape[-1]
    t_x = x_mask.reshape(b, -1).to(torch.float32).sum(1).to(torch.int32).contiguous()
    t_y = y_mask.reshape(b, -1).to(torch.float32).sum(1).to(torch.int32).contiguous()
    path = torch.empty((b, tx, ty), dtype=torch.float32, device=z_p.device)
    logp = torch.empty((b, tx, ty), dtype=torch.float32, device=z_p.device) if return_logp else None
    L = _lib.lib()
    with torch.cuda.device(z_p.device):
        ws = _lib.workspace(z_p.device, L.b200tts_mas_from_stats_workspace_bytes(b, tx, ty), "mas_from_stats")
        rc = L.b200tts_mas_from_stats(_lib.ptr(z_p), _lib.ptr(m_p), _lib.ptr(logs_p), _lib.ptr(t_x), _lib.ptr(t_y), b, c, tx,
                                      ty, _lib.ptr(path), 1, _lib.ptr(logp), _lib.ptr(ws), ctypes.c_size_t(ws.numel()),
                                      _lib.stream_ptr(z_p.device))
    _lib.check(rc, "mas_from_stats")
    attn = path.unsqueeze(1)
    return (attn, logp) if return_logp else attn
