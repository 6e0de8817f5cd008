"""CPU, world_size = 2 over gloo: ray/frame sharding is a partition, and the one-bucket gradient
all-reduce equals the single-process sum (rays shard with no data-path collective; one flat
all-reduce of the MLP gradients per training step)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recmv_b200.render import allreduce_grads, shard_frames, shard_rows


def test_shard_rows_is_a_partition():
    for H in (512, 1024, 7, 1):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                r0, n = shard_rows(H, r, world)
                covered += list(range(r0, r0 + n))
            assert covered == list(range(H))
            sizes = [shard_rows(H, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert sorted(sum((shard_frames(10, r, 4) for r in range(4)), [])) == list(range(10))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(39, 64), torch.nn.Softplus(beta=100), torch.nn.Linear(64, 1))
    x = torch.randn(64, 39, generator=torch.Generator().manual_seed(5))
    r0, n = shard_rows(64, rank, world)  # each rank owns a contiguous block of "rays"
    loss = net(x[r0:r0 + n]).sum() / 64.0   # partial loss normalised by the GLOBAL ray count
    loss.backward()
    allreduce_grads(list(net.parameters())).wait()
    q.put((rank, [p.grad.detach().numpy().copy() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_grads_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(39, 64), torch.nn.Softplus(beta=100), torch.nn.Linear(64, 1))
    x = torch.randn(64, 39, generator=torch.Generator().manual_seed(5))
    (net(x).sum() / 64.0).backward()
    for r in range(world):
        for g, p in zip(got[r], net.parameters()):
            assert torch.allclose(torch.from_numpy(g), p.grad, atol=1e-6)
