"""N>1 host logic on CPU: block partition of clips over ranks and the end-of-run gather, world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from convofusion_b200.distributed import batches, clip_seeds, gather_motions, shard_range


def test_shard_range_covers_every_unit_once():
    for n in (0, 1, 7, 64, 4096, 4099):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                s, e = shard_range(n, r, world)
                assert 0 <= s <= e <= n
                got += list(range(s, e))
            assert got == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert batches(10, 150, 64) == [(10, 74), (74, 138), (138, 150)]
    assert clip_seeds(3, 6) == [1237, 1238, 1239]


def _worker(rank, world, port, n_units, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(n_units, rank, world)
    # stand-in for the sampler output of this rank's clips: value encodes the global clip id
    local = torch.arange(s, e, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 3).contiguous()
    out = gather_motions(local, n_units)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_over_gloo_world2():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_units = 7    # ragged: rank 0 has 4 clips, rank 1 has 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.shape == (n_units, 4, 3)
    assert out[:, 0, 0].tolist() == [float(i) for i in range(n_units)]
