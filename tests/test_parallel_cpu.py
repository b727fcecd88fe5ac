"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo processes reduce the flat
gradient arena bucket by bucket and end up with identical, correctly averaged gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from keypoints_interpolation_transformer_b200 import parallel
    from keypoints_interpolation_transformer_b200.engine import ModelLayout
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    layout = ModelLayout(108, 64, 2, 4)
    n = layout.trainable
    g = torch.Generator().manual_seed(100 + rank)
    grads = torch.randn(n, generator=g)
    mine = grads.clone()
    red = parallel.BucketReducer(grads, layout.buckets)
    red.begin()
    for b in range(len(layout.buckets)):          # the order backward completes them
        red.bucket_ready(b)
    red.finish()
    others = [torch.randn(n, generator=torch.Generator().manual_seed(100 + k)) for k in range(world)]
    want = sum(others)
    ok = torch.allclose(grads, want, rtol=1e-6, atol=1e-6)
    lo, hi = parallel.shard_batch(10, rank, world)
    q.put((rank, ok, red.bytes_reduced == 4 * n, (lo, hi), float((mine - others[rank]).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] and r[2] for r in res)
    assert res[0][3] == (0, 5) and res[1][3] == (5, 10)
    assert all(r[4] == 0.0 for r in res)
