"""CPU, world_size 2, gloo: host-side logic of the multi-GPU paths -- index-range sharding + score gather, and the
bucketed gradient sink (layout, bucket completion order, averaging)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from btsbot_b200.parallel import GradSink, score_alerts, shard_range


def test_shard_ranges_tile_the_index_space():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        trip = np.arange(n, dtype=np.float32).reshape(n, 1, 1, 1) * np.ones((1, 63, 63, 3), np.float32)
        meta = np.arange(n, dtype=np.float32).reshape(n, 1) * np.ones((1, 25), np.float32)
        calls = []

        def fake_score(t, m):                     # stands in for AlertScorer: score = f(alert index) only
            calls.append(len(t))
            return torch.from_numpy(t[:, 0, 0, 0] * 2.0 + m[:, 3])

        scores = score_alerts(fake_score, trip, meta, batch_size=5, gather=True)
        lo, hi = shard_range(n, rank, world)
        assert sum(calls) == hi - lo and max(calls) <= 5
        assert torch.equal(scores, torch.arange(n, dtype=torch.float32) * 3.0)
        local = score_alerts(fake_score, trip, meta, batch_size=64, gather=False)
        assert torch.equal(local, torch.arange(lo, hi, dtype=torch.float32) * 3.0)

        # gradient sink: 3 "layers", tiny buckets; backward order = reverse parameter order
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(s)) for s in ((4, 3), (5,), (2, 2, 2))]
        sink = GradSink(params, bucket_bytes=24)
        assert sink.total == 12 + 5 + 8 and sink.offset[id(params[2])] == 0 and sink.offset[id(params[0])] == 13
        grads = [torch.full_like(p, float(rank + 1) * (i + 1)) for i, p in enumerate(params)]
        for i in (2, 1, 0):
            params[i].grad = sink.adopt(params[i], grads[i])
            sink.ready(params[i])
        sink.flush()
        for i, p in enumerate(params):
            assert torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))        # mean over ranks 1,2
        assert sorted(sink.last_order) == list(range(len(sink.bounds))) and sink.last_early == len(sink.bounds)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_and_gradient_buckets():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 23, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
