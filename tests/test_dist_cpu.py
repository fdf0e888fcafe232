"""world_size-2 gloo tests (CPU) of the scene sharding and the embedding all-gather."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("3dvlp_b200.dist")
    try:
        batch = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)
        mine = d.shard_scenes(batch)
        s, e = d.shard_range(7, world, rank)
        assert torch.equal(mine, batch[s:e])
        # equal-shape gather
        t = torch.full((2, 4), float(rank))
        g = d.concat_all_gather(t)
        assert g.shape == (2 * world, 4) and torch.equal(g[2 * rank:2 * rank + 2], t)
        assert not g.requires_grad
        # gather that keeps the local slice differentiable
        x = torch.full((3, 2), float(rank + 1), requires_grad=True)
        y = d.all_gather_with_local_grad(x)
        y.sum().backward()
        assert torch.equal(x.grad, torch.ones_like(x))
        # both embedding sets in one collective
        obj = torch.randn(5, 8, generator=torch.Generator().manual_seed(rank), requires_grad=True)
        sent = torch.randn(2, 8, generator=torch.Generator().manual_seed(10 + rank))
        o, s_ = d.gather_contrastive_embeddings(obj, sent)
        assert o.shape == (5 * world, 8) and s_.shape == (2 * world, 8)
        torch.testing.assert_close(o.norm(dim=-1), torch.ones(5 * world))
        for r in range(world):
            ref = torch.nn.functional.normalize(torch.randn(5, 8, generator=torch.Generator().manual_seed(r)), dim=-1)
            torch.testing.assert_close(o[5 * r:5 * r + 5].detach(), ref)
        (o * o).sum().backward()
        assert obj.grad is not None
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_everything():
    d = importlib.import_module("3dvlp_b200.dist")
    for total in (0, 1, 7, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [d.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
