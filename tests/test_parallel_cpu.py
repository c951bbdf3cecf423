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
        return {"model_outputs": wav, "y_lengths": y_lengths, "y_mask": torch.ones(x.shape[0], 1, t)}


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
