"""Host-side multi-GPU logic on CPU: world_size 2, gloo backend (the N>1 path of bench.py / parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tts_b200.parallel import gather_waveforms, shard_by_cost, synthesize_sharded


def test_shard_by_cost_is_a_balanced_partition():
    costs = [64, 10, 50, 50, 7, 30, 64, 1]
    shards = shard_by_cost(costs, 3)
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(len(costs)))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(costs)
    assert shard_by_cost([], 2) == [[], []]


class _FakeVits(torch.nn.Module):
    """Deterministic stand-in with the Vits.inference contract (no CUDA needed)."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))

    def inference(self, x, aux_input, **kw):
        lens = aux_input["x_lengths"]
        y_lengths = lens * 2
        t = int(y_lengths.max())
        wav = torch.zeros(x.shape[0], 1, t * 4)
        for b in range(x.shape[0]):
            n = int(y_lengths[b]) * 4
            wav[b, 0, :n] = float(x[b, : int(lens[b])].sum()) + torch.arange(n)
        return {"model_outputs": wav, "y_lengths": y_lengths, "y_mask": torch.ones(x.shape[0], 1, t),
                "wav_lengths": y_lengths * 4}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ragged gather: rank r holds r+1 utterances of different padded lengths
        wav = torch.full((rank + 1, 1, 5 + 3 * rank), float(rank + 1))
        lens = torch.arange(1, rank + 2)
        got = gather_waveforms(wav, lens, dst=0)
        if rank == 0:
            assert len(got) == world
            for r, (w, l) in enumerate(got):
                assert w.shape[0] == r + 1 and torch.all(w[:, :, : 5 + 3 * r] == r + 1) and l.tolist() == list(range(1, r + 2))
        else:
            assert got is None
        torch.manual_seed(0)
        tokens = torch.randint(1, 9, (5, 12))
        x_lengths = torch.tensor([12, 3, 7, 12, 5])
        res = synthesize_sharded(_FakeVits(), tokens, x_lengths, dst=0)
        if rank == 0:
            for i in range(5):
                n = int(x_lengths[i]) * 8
                want = float(tokens[i, : int(x_lengths[i])].sum()) + torch.arange(n)
                assert torch.equal(res[i], want), i
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gather_and_sharded_synthesis_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_bucket_by_length_partitions_within_budget():
    from tts_b200.parallel import bucket_by_length
    lens = [5, 64, 33, 64, 1, 17, 40, 40, 8, 64, 2]
    buckets = bucket_by_length(lens, max_padded_tokens=128, max_batch=4)
    assert sorted(i for b in buckets for i in b) == list(range(len(lens)))
    for b in buckets:
        assert len(b) <= 4 and len(b) * max(lens[i] for i in b) <= 128
    # similar lengths end up together: the three 64-token items cannot share a 128-token budget beyond two
    assert all(len(b) <= 2 for b in buckets if max(lens[i] for i in b) == 64)
    assert bucket_by_length([], 128, 4) == []
    assert bucket_by_length([500], 128, 4) == [[0]]          # one over-budget item still gets its own batch


def test_synthesize_batched_matches_one_by_one_and_keeps_order():
    from tts_b200.parallel import concat_sentences, synthesize_batched, to_int16
    torch.manual_seed(0)
    model = _FakeVits()
    seqs = [torch.randint(1, 50, (n,)).tolist() for n in (7, 31, 3, 18, 31, 1, 12)]
    one_by_one = [model.inference(torch.tensor([s]), {"x_lengths": torch.tensor([len(s)])})["model_outputs"][0, 0]
                  for s in seqs]
    got = synthesize_batched(model, seqs, max_padded_tokens=64, max_batch=3)
    assert len(got) == len(seqs)
    for a, b in zip(got, one_by_one):
        assert torch.equal(a, b)
    speakers = torch.arange(len(seqs))

    class _Spk(_FakeVits):
        def inference(self, x, aux_input, **kw):
            out = super().inference(x, aux_input, **kw)
            out["model_outputs"] = out["model_outputs"] + 1000.0 * aux_input["speaker_ids"].view(-1, 1, 1).float()
            return out

    got = synthesize_batched(_Spk(), seqs, {"speaker_ids": speakers}, max_padded_tokens=64, max_batch=3)
    for i, (a, b) in enumerate(zip(got, one_by_one)):
        assert torch.equal(a, b + 1000.0 * i)                  # conditioning follows its sentence through the buckets
    cat = concat_sentences(got[:2], gap=10)
    assert cat.numel() == got[0].numel() + got[1].numel() + 20 and float(cat[got[0].numel():got[0].numel() + 10].abs().sum()) == 0
    import numpy as np
    w = torch.randn(1000) * 0.3
    want = (w.numpy() * (32767 / max(0.01, np.max(np.abs(w.numpy()))))).astype(np.int16)
    assert np.array_equal(to_int16(w).numpy(), want)
    assert to_int16(torch.zeros(4)).tolist() == [0, 0, 0, 0]
