"""Data-parallel batched inference: utterances are independent (no cross-utterance op anywhere on the path,
SURVEY section 8e), so the batch is dealt over ranks by predicted cost and the only collective is the final gather
of waveforms + lengths to one rank (NCCL over NVLink on the GPU box; the same code runs on gloo for the CPU tests).
The reference has no multi-GPU inference at all (replicas = separate processes)."""
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_by_cost(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first deal: returns, per rank, the indices of the items it owns.
    Cost is proportional to decoder frames (HiFiGAN dominates), for which text length is the available proxy."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += float(costs[i])
    return [sorted(s) for s in shards]


def gather_waveforms(wav: torch.Tensor, lengths: torch.Tensor, dst: int = 0, group=None):
    """wav [b_r, 1, T_r] (per-rank padded length), lengths int64 [b_r] (valid samples or frames).
    Returns on ``dst`` a list (one entry per rank) of (wav [b_r,1,T_max], lengths [b_r]); None elsewhere.
    Ranks may hold different batch sizes and different T."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    meta = torch.tensor([wav.shape[0], wav.shape[-1]], dtype=torch.int64, device=wav.device)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    bmax = max(int(m[0]) for m in metas)
    tmax = max(int(m[1]) for m in metas)
    pad = torch.zeros((bmax, 1, tmax), dtype=wav.dtype, device=wav.device)
    pad[: wav.shape[0], :, : wav.shape[-1]] = wav
    lpad = torch.zeros((bmax,), dtype=torch.int64, device=wav.device)
    lpad[: lengths.shape[0]] = lengths
    outs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    louts = [torch.empty_like(lpad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, outs, dst=dst, group=group)
    dist.gather(lpad, louts, dst=dst, group=group)
    if rank != dst:
        return None
    return [(outs[r][: int(metas[r][0])], louts[r][: int(metas[r][0])]) for r in range(world)]


def synthesize_sharded(model, tokens: torch.Tensor, x_lengths: torch.Tensor, aux_input=None, dst: int = 0, **kw):
    """Every rank passes the SAME full batch (host tensors); each synthesises its shard with ``model.inference``
    and ``dst`` receives all waveforms re-ordered to the input order: list of 1-D tensors (valid samples only)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    shards = shard_by_cost([float(v) for v in x_lengths.tolist()], world)
    mine = shards[rank]
    dev = next(model.parameters()).device
    aux = dict(aux_input or {})
    idx = torch.tensor(mine, dtype=torch.long)
    tmax = int(x_lengths[idx].max()) if mine else 1
    aux["x_lengths"] = x_lengths[idx].to(dev)
    for k in ("speaker_ids", "d_vectors", "language_ids"):
        if aux.get(k, None) is not None:
            aux[k] = aux[k][idx].to(dev)
    out = model.inference(tokens[idx, :tmax].to(dev), aux, **kw)
    hop = out["model_outputs"].shape[-1] // max(out["y_mask"].shape[-1], 1)
    got = gather_waveforms(out["model_outputs"], out["y_lengths"] * hop, dst=dst)
    if got is None:
        return None
    result = [None] * tokens.shape[0]
    for r, (wav, lens) in enumerate(got):
        for j, i in enumerate(shards[r]):
            result[i] = wav[j, 0, : int(lens[j])]
    return result


# ----------------------------------------------------------------------------- sentence-level batching
def bucket_by_length(lengths: Sequence[int], max_padded_tokens: int = 4096, max_batch: int = 64) -> List[List[int]]:
    """Groups item indices into batches of similar length: items are taken longest first and a batch is closed when
    adding the next one would exceed ``max_batch`` items or ``max_padded_tokens`` = items x longest item (the padded
    work the text encoder and, through the durations, the decoder will do).  Every index appears exactly once."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    buckets: List[List[int]] = []
    cur: List[int] = []
    cur_max = 0
    for i in order:
        n = max(int(lengths[i]), 1)
        new_max = max(cur_max, n)
        if cur and (len(cur) + 1 > max_batch or (len(cur) + 1) * new_max > max_padded_tokens):
            buckets.append(cur)
            cur, new_max = [], n
        cur.append(i)
        cur_max = new_max
    if cur:
        buckets.append(cur)
    return buckets


def synthesize_batched(model, token_seqs: Sequence[Sequence[int]], aux_input=None, pad_id: int = 0,
                       max_padded_tokens: int = 4096, max_batch: int = 64, **kw) -> List[torch.Tensor]:
    """What ``Synthesizer.tts`` does one sentence at a time (TTS/utils/synthesizer.py:384-441: tokenise a sentence,
    ``model.inference`` with batch 1, append), done as a few padded batches: ``token_seqs`` are the already
    tokenised sentences, per-sentence conditioning in ``aux_input`` (``speaker_ids`` / ``d_vectors`` /
    ``language_ids``, first dimension = sentence) is carried along, and the result is one 1-D waveform (valid
    samples only) per sentence in the input order.  Utterances never interact on the path (masks are per row), so
    a sentence's audio does not depend on what it was batched with (given the same random draws)."""
    n = len(token_seqs)
    lengths = [len(s) for s in token_seqs]
    if any(l == 0 for l in lengths):
        raise ValueError("tts_b200.synthesize_batched: empty token sequence")
    dev = next(model.parameters()).device
    aux_all = dict(aux_input or {})
    result: List[torch.Tensor] = [None] * n
    for bucket in bucket_by_length(lengths, max_padded_tokens, max_batch):
        tmax = max(lengths[i] for i in bucket)
        tok = torch.full((len(bucket), tmax), int(pad_id), dtype=torch.int64)
        for j, i in enumerate(bucket):
            tok[j, : lengths[i]] = torch.as_tensor(list(token_seqs[i]), dtype=torch.int64)
        aux = {"x_lengths": torch.tensor([lengths[i] for i in bucket], dtype=torch.int64, device=dev)}
        idx = torch.tensor(bucket, dtype=torch.long)
        for k in ("speaker_ids", "d_vectors", "language_ids"):
            if aux_all.get(k, None) is not None:
                aux[k] = aux_all[k][idx].to(dev)
        out = model.inference(tok.to(dev), aux, **kw)
        hop = out["model_outputs"].shape[-1] // max(out["y_mask"].shape[-1], 1)
        valid = (out["y_lengths"] * hop).tolist()
        for j, i in enumerate(bucket):
            result[i] = out["model_outputs"][j, 0, : int(valid[j])]
    return result


def concat_sentences(wavs: Sequence[torch.Tensor], gap: int = 10000) -> torch.Tensor:
    """``wavs += list(waveform); wavs += [0] * 10000`` (synthesizer.py:440-441): sentences joined with 10000 zero
    samples after each one."""
    parts = []
    for w in wavs:
        parts.append(w.reshape(-1))
        parts.append(torch.zeros(gap, dtype=w.dtype, device=w.device))
    return torch.cat(parts) if parts else torch.zeros(0)


def to_int16(wav: torch.Tensor) -> torch.Tensor:
    """The peak normalisation ``save_wav`` applies before writing (TTS/utils/audio/numpy_transforms.py:438-440):
    ``wav * (32767 / max(0.01, max|wav|))`` truncated to int16."""
    peak = float(wav.abs().max()) if wav.numel() else 0.0
    return (wav * (32767.0 / max(0.01, peak))).to(torch.int16)
