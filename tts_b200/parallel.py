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


class WaveformGather:
    """The one collective of the data-parallel path: every rank's padded waveforms + valid lengths to ``dst``.

    What the first version paid for 6 MB per rank (+7..10 ms per step at 2..8 GPUs) was not the transfer but its
    set-up: a metadata all-gather followed by 2*N device->host reads, freshly allocated pad / receive buffers and
    three blocking collectives on the compute stream.  Here
      * shapes travel as HOST integers ((b, T) are known to the host when ``Vits.inference`` returns) through a
        pinned 2-word all-gather on a side stream, so the host never waits for the compute stream;
      * lengths ride in the same message as the samples (one uint8 payload = [b_max*T_max fp32 | b_max int64]);
      * send / receive buffers persist and only grow; the collective runs on the side stream after an event on the
        compute stream, so the next step's kernels may overlap it; ``record_stream`` keeps the waveform alive.
    ``gather`` returns at once; the result's tensors are valid after ``result.wait()`` (makes the current stream
    wait for the side stream) -- ``None`` on ranks other than ``dst``.  On gloo (CPU tests) the same code runs
    without streams."""

    class Result:
        def __init__(self, parts, event):
            self.parts, self._event = parts, event

        def wait(self):
            if self._event is not None:
                torch.cuda.current_stream().wait_event(self._event)
            return self.parts

    def __init__(self, device, dst: int = 0, group=None):
        self.device, self.dst, self.group = torch.device(device), dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cuda = self.device.type == "cuda"
        self.side = torch.cuda.Stream(self.device) if self.cuda else None
        self.meta_host = torch.zeros(2, dtype=torch.int64)
        self.metas_host = torch.zeros(2 * self.world, dtype=torch.int64)
        if self.cuda:
            self.meta_host, self.metas_host = self.meta_host.pin_memory(), self.metas_host.pin_memory()
        self.meta_dev = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.metas_dev = torch.zeros(2 * self.world, dtype=torch.int64, device=self.device)
        self.send = None
        self.recv = None

    def _exchange_shapes(self, b, t):
        self.meta_host[0], self.meta_host[1] = int(b), int(t)
        if not self.cuda:
            dist.all_gather_into_tensor(self.metas_host, self.meta_host, group=self.group)
            return self.metas_host.view(self.world, 2).tolist()
        with torch.cuda.stream(self.side):
            self.meta_dev.copy_(self.meta_host, non_blocking=True)
            dist.all_gather_into_tensor(self.metas_dev, self.meta_dev, group=self.group)
            self.metas_host.copy_(self.metas_dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.side)
        ev.synchronize()      # side stream only: the compute stream keeps running
        return self.metas_host.view(self.world, 2).tolist()

    def gather(self, wav: torch.Tensor, lengths: torch.Tensor):
        """wav [b_r, 1, T_r] fp32 (per-rank batch and padded length), lengths int64 [b_r] (valid samples)."""
        b, t = wav.shape[0], wav.shape[-1]
        metas = self._exchange_shapes(b, t)
        bmax, tmax = max(m[0] for m in metas), max(m[1] for m in metas)
        wav_bytes, nbytes = bmax * tmax * 4, bmax * tmax * 4 + bmax * 8
        if self.send is None or self.send.numel() < nbytes:
            cap = int(nbytes * 1.25) // 16 * 16 + 16
            self.send = torch.zeros(cap, dtype=torch.uint8, device=self.device)
            self.recv = torch.zeros(self.world * cap, dtype=torch.uint8, device=self.device) if self.rank == self.dst else None
        main_done = None
        if self.cuda:
            main_done = torch.cuda.Event()
            main_done.record(torch.cuda.current_stream(self.device))
        ctx = torch.cuda.stream(self.side) if self.cuda else _NullCtx()
        with ctx:
            if self.cuda:
                self.side.wait_event(main_done)
                wav.record_stream(self.side)
                lengths.record_stream(self.side)
            payload = self.send[:nbytes]
            pw = payload[:wav_bytes].view(torch.float32).view(bmax, tmax)
            if t < tmax or b < bmax:
                pw.zero_()
            pw[:b, :t].copy_(wav.reshape(b, t))
            pl = payload[wav_bytes:].view(torch.int64)
            pl.zero_()
            pl[:b].copy_(lengths.to(torch.int64))
            outs = None
            if self.rank == self.dst:
                outs = [self.recv[r * nbytes:(r + 1) * nbytes] for r in range(self.world)]
            dist.gather(payload, outs, dst=self.dst, group=self.group)
            done = None
            if self.cuda:
                done = torch.cuda.Event()
                done.record(self.side)
        if self.rank != self.dst:
            return WaveformGather.Result(None, done)
        parts = []
        for r in range(self.world):
            br = metas[r][0]
            w = outs[r][:wav_bytes].view(torch.float32).view(bmax, 1, tmax)[:br]
            parts.append((w, outs[r][wav_bytes:].view(torch.int64)[:br]))
        return WaveformGather.Result(parts, done)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_gatherers = {}


def gather_waveforms(wav: torch.Tensor, lengths: torch.Tensor, dst: int = 0, group=None):
    """wav [b_r, 1, T_r] (per-rank padded length), lengths int64 [b_r] (valid samples or frames).
    Returns on ``dst`` a list (one entry per rank) of (wav [b_r,1,T_max], lengths [b_r]); None elsewhere.
    Ranks may hold different batch sizes and different T.  (Blocking convenience form of ``WaveformGather``: the
    returned tensors are views of its persistent receive buffer, valid until the next call.)"""
    key = (str(wav.device), dst, id(group))
    g = _gatherers.get(key)
    if g is None:
        g = _gatherers[key] = WaveformGather(wav.device, dst, group)
    return g.gather(wav, lengths).wait()


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
    got = gather_waveforms(out["model_outputs"], out["wav_lengths"], dst=dst)
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
                       max_padded_tokens: int = 4096, max_batch: int = 64, sdp_noise=None, prior_noise=None,
                       **kw) -> List[torch.Tensor]:
    """What ``Synthesizer.tts`` does one sentence at a time (TTS/utils/synthesizer.py:384-441: tokenise a sentence,
    ``model.inference`` with batch 1, append), done as a few padded batches: ``token_seqs`` are the already
    tokenised sentences, per-sentence conditioning in ``aux_input`` (``speaker_ids`` / ``d_vectors`` /
    ``language_ids``, first dimension = sentence) is carried along, and the result is one 1-D waveform (valid
    samples only) per sentence in the input order.  Utterances never interact on the path (masks are per row), so
    a sentence's audio does not depend on what it was batched with (given the same random draws) -- except, exactly as
    in the reference's own batched ``Vits.inference``, within the decoder's receptive field of the sentence's end, where
    a padded row sees the decoder's response to zero frames instead of the convolutions' zero padding.

    ``sdp_noise`` (sequence of [2, T_i] tensors) / ``prior_noise`` (callable ``(sentence_index, channels, frames) ->
    [channels, frames]``) pin the two random draws PER SENTENCE (tests compare with sentence-at-a-time synthesis)."""
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
        kwb = dict(kw)
        if sdp_noise is not None:
            nz = torch.zeros(len(bucket), 2, tmax)
            for j, i in enumerate(bucket):
                nz[j, :, : lengths[i]] = sdp_noise[i]
            kwb["sdp_noise"] = nz
        if prior_noise is not None:
            def rows(shape, y_lengths, bucket=bucket):
                out_ = torch.zeros(shape)
                for j, i in enumerate(bucket):
                    n = int(y_lengths[j])
                    out_[j, :, :n] = prior_noise(i, shape[1], n)
                return out_.to(dev)
            rows.wants_lengths = True
            kwb["prior_noise"] = rows
        out = model.inference(tok.to(dev), aux, **kwb)
        valid = out["wav_lengths"].tolist()
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
    ``wav * (32767 / max(0.01, max|wav|))`` truncated to int16.  CUDA tensors go through the device kernels
    (``b200tts_absmax`` / ``b200tts_to_int16``, no host round trip); host tensors use the same arithmetic in torch."""
    if wav.is_cuda:
        from .vocoder import wav_to_int16
        return wav_to_int16(wav)
    peak = float(wav.abs().max()) if wav.numel() else 0.0
    return (wav * (32767.0 / max(0.01, peak))).to(torch.int16)


def synthesize_to_int16(model, token_seqs: Sequence[Sequence[int]], aux_input=None, gap: int = 10000, **kw) -> torch.Tensor:
    """``Synthesizer.tts`` + ``save_wav`` for already tokenised sentences, entirely on the device: batched synthesis,
    sentences joined with ``gap`` zero samples (synthesizer.py:440-441), one peak over the whole utterance and the int16
    conversion (numpy_transforms.py:439-441).  Returns the int16 samples ``save_wav`` would write (CUDA tensor)."""
    wavs = synthesize_batched(model, token_seqs, aux_input, **kw)
    return to_int16(concat_sentences(wavs, gap))
