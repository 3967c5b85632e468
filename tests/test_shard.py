"""Multi-GPU host logic on CPU: world_size-2 gloo run of the batch sharding + final pose all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from catre_b200 import shard


def test_shard_bounds_cover_everything():
    for total in (0, 1, 3, 64, 65, 512):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, k1, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(123)
    poses = torch.randn(k1, total, 3, 4, generator=g)
    scales = torch.randn(k1, total, 3, generator=g)
    lo, hi = shard.shard_bounds(total, world, rank)
    fp, fs = shard.gather_poses(poses[:, lo:hi].contiguous(), scales[:, lo:hi].contiguous(), total)
    ok = bool(torch.equal(fp, poses) and torch.equal(fs, scales))
    # the bench's form: the engine's packed final poses [per, 15] of each rank's slice, ONE collective into a preallocated buffer
    per = shard.shard_size(total, world)
    local = torch.zeros(per, 15)
    local[: hi - lo] = shard.pack_poses(poses[-1:, lo:hi], scales[-1:, lo:hi])[0]
    out = shard.gather_packed(local, torch.empty(world * per, 15))
    ok &= bool(torch.equal(out[:total], shard.pack_poses(poses[-1:], scales[-1:])[0]))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_poses_gloo_world2():
    ctx = mp.get_context("spawn")
    for total in (7, 8):  # uneven and even splits
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, 5, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(60)
        assert res == [(0, True), (1, True)]
