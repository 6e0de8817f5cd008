"""2 GPUs over NCCL (skipped on a 1-GPU box): the sharded render needs no data-path collective -- each rank
renders its own block of image rows through the fused kernel and the concatenation equals the single-GPU frame --
and the one-bucket gradient all-reduce on a side stream equals the single-process gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    from recmv_b200 import synth
    from recmv_b200.render import SdfRenderer, allreduce_grads, shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    # (1) row-sharded render of ONE frame: replicated weights / voxel / pose, no collective on the data path
    ren = SdfRenderer(dev, seed=0, samples=16)
    poses, trans = synth.poses_trans(1, seed=11)
    A, t = ren.bone_matrices(poses.to(dev), trans.to(dev))
    H = 64
    r0, n = shard_rows(H, rank, world)
    dirs = synth.pinhole_rays(H, 64, device=dev, row0=r0, rows=n)
    sdf, _, hit_idx, hit_t = ren.render(dirs, A, t)
    # (2) gradient all-reduce of a small SDF-like net on a communication stream
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(39, 64), torch.nn.Softplus(beta=100), torch.nn.Linear(64, 1)).to(dev)
    x = torch.randn(64, 39, generator=torch.Generator().manual_seed(5)).to(dev)
    b0, bn = shard_rows(64, rank, world)
    (net(x[b0:b0 + bn]).sum() / 64.0).backward()
    allreduce_grads(list(net.parameters()), comm_stream=torch.cuda.Stream(dev)).wait()
    # (3) the real thing: data-parallel training step of the SDF network (fused forward + tcgen05 backward), each rank on
    # its shard of the points, loss normalised by the GLOBAL point count, one flat all-reduce of all parameter gradients
    from recmv_b200 import ops, testing
    from recmv_b200.model import getTmpSdf
    sdfnet = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=101).to(dev)
    P = 1024
    xs = (torch.rand((P, 3), generator=torch.Generator().manual_seed(9)) * 1.2 - 0.6).to(dev)
    s0, sn = shard_rows(P, rank, world)
    y = sdfnet(xs[s0:s0 + sn].clone().requires_grad_(False), 0.7)
    ((y.pow(2).sum() + sdfnet.rendcond.pow(2).sum() * 0.01) / P).backward()
    assert sdfnet.last_path == "fused-train" and ops.SdfMlpTrainFunction.last_backward == "fused-tcgen05"
    allreduce_grads(list(sdfnet.parameters()), comm_stream=torch.cuda.Stream(dev)).wait()
    torch.cuda.synchronize(dev)
    ops.check_async_errors()
    q.put((rank, r0, sdf.cpu().numpy(), hit_idx.cpu().numpy(), [p.grad.cpu().numpy() for p in net.parameters()],
           [p.grad.cpu().numpy() for p in sdfnet.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_render_and_grad_allreduce():
    from recmv_b200 import synth
    from recmv_b200.render import SdfRenderer
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in range(world)), key=lambda g: g[1])
    for p in procs:
        p.join(120)
    dev = torch.device("cuda", 0)
    ren = SdfRenderer(dev, seed=0, samples=16)
    poses, trans = synth.poses_trans(1, seed=11)
    A, t = ren.bone_matrices(poses.to(dev), trans.to(dev))
    sdf, _, hit_idx, _ = ren.render(synth.pinhole_rays(64, 64, device=dev), A, t)
    import numpy as np
    assert np.array_equal(np.concatenate([g[2] for g in got]), sdf.cpu().numpy())      # same kernel, same rows: bit-equal
    assert np.array_equal(np.concatenate([g[3] for g in got]), hit_idx.cpu().numpy())
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(39, 64), torch.nn.Softplus(beta=100), torch.nn.Linear(64, 1)).to(dev)
    x = torch.randn(64, 39, generator=torch.Generator().manual_seed(5)).to(dev)
    (net(x).sum() / 64.0).backward()
    for g in got:
        for a, p in zip(g[4], net.parameters()):
            assert torch.allclose(torch.from_numpy(a).to(dev), p.grad, atol=1e-6)
    # data-parallel SDF training step == the single-GPU full-batch step (up to the summation order of the all-reduce)
    from recmv_b200 import testing
    from recmv_b200.model import getTmpSdf
    sdfnet = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=101).to(dev)
    P = 1024
    xs = (torch.rand((P, 3), generator=torch.Generator().manual_seed(9)) * 1.2 - 0.6).to(dev)
    y = sdfnet(xs, 0.7)
    ((y.pow(2).sum() + sdfnet.rendcond.pow(2).sum() * 0.01) / P).backward()
    for g in got:
        for a, p in zip(g[5], sdfnet.parameters()):
            a = torch.from_numpy(a).to(dev)
            assert float((a - p.grad).abs().max()) <= 2e-5 * float(p.grad.abs().max()) + 1e-12
        for a, b in zip(g[5], got[0][5]):
            assert np.array_equal(a, b)                     # every rank holds the same reduced gradients
