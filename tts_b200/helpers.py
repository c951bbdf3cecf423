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
