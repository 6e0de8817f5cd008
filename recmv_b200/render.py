"""The fused render path as a user-facing object, plus its data-parallel sharding.

`SdfRenderer` owns the device-resident state of one scene (packed SDF weights, channels-last skinning
voxel, skeleton) and exposes

    render(ray_dirs, poses, trans)           device tensors in, device tensors out
    render_host(ray_dirs_pinned, poses, ...) HOST buffers in, HOST result out (H2D + kernels + D2H) --
                                             the call bench.py times as `e2e`

Sharding (`shard_rows`): rays are independent given replicated weights / voxel / bone matrices
(SURVEY 8e), so ranks take contiguous blocks of image rows (or whole frames) and there is NO
data-path collective; gradients -- when training -- are exchanged by `allreduce_grads`.
"""
import torch

from . import ops, synth
from .model import LBSkinner, getTmpSdf


def shard_rows(height, rank, world):
    """Contiguous block of image rows for `rank` (first `height % world` ranks get one extra row)."""
    base, extra = divmod(height, world)
    row0 = rank * base + min(rank, extra)
    return row0, base + (1 if rank < extra else 0)


def shard_frames(num_frames, rank, world):
    """Frames a rank owns when there are at least as many frames as ranks."""
    return list(range(rank, num_frames, world))


def allreduce_grads(params, group=None, comm_stream=None):
    """ONE flat fp32 all-reduce (sum) over the gradients of `params` (SDF nets + translator + rendnet;
    deformer + pose only in large-pose mode), issued on `comm_stream` so it overlaps the caller's
    remaining backward work.  Returns a handle whose .wait() scatters the reduced bucket back."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        class _Done:
            def wait(self):
                return None
        return _Done()
    dev = grads[0].device
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    if comm_stream is not None and dev.type == "cuda":
        comm_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(comm_stream):
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
    else:
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)

    class _Handle:
        def wait(self_inner):
            work.wait()
            if comm_stream is not None and dev.type == "cuda":
                torch.cuda.current_stream(dev).wait_stream(comm_stream)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
    return _Handle()


class SdfRenderer:
    def __init__(self, device, sdf_net=None, voxel_shape=(65, 225, 129), seed=0, mode=None,
                 samples=64, t_near=synth.T_NEAR, t_far=synth.T_FAR, cam_pos=synth.CAM_POS):
        self.device = torch.device(device)
        self.mode = ops.DEFAULT_MLP_MODE if mode is None else mode
        self.samples, self.t_near, self.t_far, self.cam_pos = samples, t_near, t_far, cam_pos
        if sdf_net is None:
            torch.manual_seed(seed)
            sdf_net = getTmpSdf(self.device, 6, 0.6, 256)
        self.sdf_net = sdf_net
        Js, parents, init = synth.skeleton()
        ws = synth.skinning_voxel(voxel_shape, seed=7, device=self.device)
        self.skinner = LBSkinner(ws, [-1.1] * 3, [1.1] * 3, Js, parents, init_pose=init,
                                 bbox_extend=torch.tensor(synth.BBOX_EXTEND),
                                 bbox_center=torch.tensor(synth.BBOX_CENTER)).to(self.device)
        self.ws_cl = self.skinner.ws_channels_last()
        self.packed = sdf_net.packed_weights()
        self.pe_w = [1.0] * 12
        self._sdf_buf = None

    def bone_matrices(self, poses, trans):
        with torch.no_grad():
            A = self.skinner.bone_matrices(poses)
            t = trans + self.skinner.extra_trans
        return A.contiguous(), t.contiguous()

    def render(self, ray_dirs, A, trans, rays_per_frame=0, want_xc=False):
        R = ray_dirs.shape[0]
        if self._sdf_buf is None or self._sdf_buf.shape[0] != R:
            self._sdf_buf = torch.empty((R, self.samples), dtype=torch.float32, device=self.device)
        return ops.render_sdf(ray_dirs, self.cam_pos, self.t_near, self.t_far, self.samples, A, trans,
                              self.ws_cl, synth.BBOX_CENTER, synth.BBOX_EXTEND, self.packed, self.pe_w,
                              self.mode, None, rays_per_frame or R, want_xc, True, self._sdf_buf)

    def render_host(self, ray_dirs_pinned, A_pinned, trans_pinned, out_hit_t_pinned, out_hit_idx_pinned):
        """End-to-end call on HOST buffers (pinned): H2D inputs, fused render, D2H per-ray result."""
        d = ray_dirs_pinned.to(self.device, non_blocking=True)
        A = A_pinned.to(self.device, non_blocking=True)
        t = trans_pinned.to(self.device, non_blocking=True)
        _, _, hit_idx, hit_t = self.render(d, A, t)
        out_hit_t_pinned.copy_(hit_t, non_blocking=True)
        out_hit_idx_pinned.copy_(hit_idx, non_blocking=True)
        return out_hit_t_pinned, out_hit_idx_pinned
