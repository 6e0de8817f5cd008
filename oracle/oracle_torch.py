"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch / numpy, fp32 or fp64) of the
reference's native ops and of the composed hot path.  Nothing under recmv_b200/ may import
this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.

Every function cites the reference file:line (relative to /root/reference) it restates.
Pinned against the reference's own checks in tests/test_oracle_pins.py:
  * grid_sample3d_fwd == F.grid_sample(bilinear, border, align_corners=False)
    (MCAcc/check_grid_sampler_mine.py:8-9) and gradcheck 1st/2nd order (ibid. 10-15)
  * minv3x3: inv @ m == I on randn(10000,3,3) (FastMinv/check.py:18-20)
  * MLPs / PE: equality with the imported reference modules (oracle/refload.py) -> tests/golden
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# FastMinv  (FastMinv/Matrix3x3InvKernels.cu:21-61 forward, :63-104 backward)
# ----------------------------------------------------------------------------------------------
def minv3x3_fwd(ms: torch.Tensor):
    """Cofactor inverse; |det| < 1e-4 -> inverse = 0, check = False."""
    m = ms.reshape(-1, 3, 3)
    c00 = m[:, 1, 1] * m[:, 2, 2] - m[:, 1, 2] * m[:, 2, 1]
    c01 = -m[:, 1, 0] * m[:, 2, 2] + m[:, 1, 2] * m[:, 2, 0]
    c02 = m[:, 1, 0] * m[:, 2, 1] - m[:, 1, 1] * m[:, 2, 0]
    c10 = -m[:, 0, 1] * m[:, 2, 2] + m[:, 0, 2] * m[:, 2, 1]
    c11 = m[:, 0, 0] * m[:, 2, 2] - m[:, 0, 2] * m[:, 2, 0]
    c12 = -m[:, 0, 0] * m[:, 2, 1] + m[:, 0, 1] * m[:, 2, 0]
    c20 = m[:, 0, 1] * m[:, 1, 2] - m[:, 0, 2] * m[:, 1, 1]
    c21 = -m[:, 0, 0] * m[:, 1, 2] + m[:, 0, 2] * m[:, 1, 0]
    c22 = m[:, 0, 0] * m[:, 1, 1] - m[:, 0, 1] * m[:, 1, 0]
    det = m[:, 0, 0] * c00 + m[:, 0, 1] * c01 + m[:, 0, 2] * c02
    ok = ~(det.abs() < 0.0001)
    safe = torch.where(ok, det, torch.ones_like(det))
    inv = torch.stack([c00, c10, c20, c01, c11, c21, c02, c12, c22], dim=1) / safe[:, None]
    inv = torch.where(ok[:, None], inv, torch.zeros_like(inv)).reshape(-1, 3, 3)
    return inv, ok


def minv3x3_bwd(grads: torch.Tensor, invs: torch.Tensor):
    """out = -inv^T g inv^T  (expanded at Matrix3x3InvKernels.cu:91-102)."""
    it = invs.reshape(-1, 3, 3).transpose(1, 2)
    return -(it @ grads.reshape(-1, 3, 3) @ it)


# ----------------------------------------------------------------------------------------------
# GridSamplerMine  (MCAcc/cuda/GridSamplerMineKernel.cu:160-328 fwd, 331-570 bwd, 573-914 bwd2)
# ----------------------------------------------------------------------------------------------
def _unnormalize_clip(g, size):
    # ((x+1)*W-1)/2 (GridSamplerMineKernel.cu:210-212) then clip_coordinates_set_grad (:42-59):
    # the derivative is 0 when the unclipped coordinate is <= 0 or >= size-1.
    x = ((g + 1.0) * size - 1.0) / 2.0
    inside = (x > 0) & (x < size - 1)
    xc = x.clamp(0, size - 1)
    return torch.where(inside, x, xc.detach())


def grid_sample3d_fwd(inp: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """Trilinear, border padding, align_corners=False.  inp [N,C,D,H,W], grid [N,Do,Ho,Wo,3]
    (x->W, y->H, z->D).  Differentiable to any order in both arguments (pure torch ops)."""
    N, C, D, H, W = inp.shape
    _, Do, Ho, Wo, _ = grid.shape
    P = Do * Ho * Wo
    g = grid.reshape(N, P, 3)
    ix = _unnormalize_clip(g[..., 0], W)
    iy = _unnormalize_clip(g[..., 1], H)
    iz = _unnormalize_clip(g[..., 2], D)
    x0 = torch.floor(ix.detach())
    y0 = torch.floor(iy.detach())
    z0 = torch.floor(iz.detach())
    fx, fy, fz = ix - x0, iy - y0, iz - z0
    x0 = x0.long(); y0 = y0.long(); z0 = z0.long()
    flat = inp.reshape(N, C, D * H * W)
    out = 0
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                wgt = (fx if dx else 1 - fx) * (fy if dy else 1 - fy) * (fz if dz else 1 - fz)
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H) & (zi >= 0) & (zi < D)
                lin = (zi.clamp(0, D - 1) * H + yi.clamp(0, H - 1)) * W + xi.clamp(0, W - 1)
                v = torch.gather(flat, 2, lin[:, None, :].expand(N, C, P))
                out = out + v * (wgt * ok.to(wgt.dtype))[:, None, :]
    return out.reshape(N, C, Do, Ho, Wo)


def grid_sample3d_bwd(inp, grid, grad_out):
    """(grad_input, grad_grid) of grid_sample3d_fwd -- differentiable again (create_graph)."""
    with torch.enable_grad():
        i = inp if inp.requires_grad else inp.detach().requires_grad_(True)
        g = grid if grid.requires_grad else grid.detach().requires_grad_(True)
        out = grid_sample3d_fwd(i, g)
        gi, gg = torch.autograd.grad(out, (i, g), grad_out, create_graph=True, allow_unused=True)
    if gi is None:
        gi = torch.zeros_like(inp)
    if gg is None:
        gg = torch.zeros_like(grid)
    return gi, gg


def grid_sample3d_bwd2(ggi, ggg, inp, grid, grad_out):
    """VJP of grid_sample3d_bwd: cotangents (ggi ~ input, ggg ~ grid) -> (gI, gG, ggO)."""
    with torch.enable_grad():
        i = inp.detach().requires_grad_(True)
        g = grid.detach().requires_grad_(True)
        go = grad_out.detach().requires_grad_(True)
        gi, gg = grid_sample3d_bwd(i, g, go)
        outs = torch.autograd.grad((gi, gg), (i, g, go), (ggi, ggg), allow_unused=True)
    outs = [o if o is not None else torch.zeros_like(t) for o, t in zip(outs, (inp, grid, grad_out))]
    return tuple(outs)


# ----------------------------------------------------------------------------------------------
# smpl_pytorch.util.batch_rodrigues (un-vendored; standard HMR form; parity UNPINNED)
# ----------------------------------------------------------------------------------------------
def quat2mat(quat):  # utils/utils.py:21-38
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def batch_rodrigues(theta):
    l1 = torch.norm(theta + 1e-8, p=2, dim=1)
    angle = l1.unsqueeze(-1)
    n = theta / angle
    angle = angle * 0.5
    quat = torch.cat([torch.cos(angle), torch.sin(angle) * n], dim=1)
    return quat2mat(quat)


def scatter(src, index, dim=0, out=None, reduce="sum"):
    """torch_scatter.scatter subset used at utils/FindSurfacePs.py:30 (min, with out)."""
    red = {"min": "amin", "max": "amax", "sum": "sum", "mean": "mean"}[reduce]
    if out is None:
        size = int(index.max().item()) + 1 if index.numel() else 0
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
        return out.scatter_reduce(dim, index, src, red, include_self=False)
    return out.scatter_reduce(dim, index, src, red, include_self=True)


# ----------------------------------------------------------------------------------------------
# Embedder / annealing weights  (model/Embedder.py:43-50, utils/utils.py:40-46)
# ----------------------------------------------------------------------------------------------
def annealing_weights(multires, ratio):
    """ratio None -> all ones; ratio <= 0 -> zeros (model/network.py:93-99)."""
    if ratio is None:
        return [1.0] * (2 * multires)
    if ratio <= 0:
        return [0.0] * (2 * multires)
    alpha = ratio * multires
    out = []
    for k in range(multires):
        w = (1.0 - np.cos(np.pi * min(max(alpha - float(k), 0.0), 1.0))) / 2.0
        out.extend([w, w])
    return out


def embed(x, multires, ws=None):
    """[x, w0 sin x, w0 cos x, ..., sin 2^(L-1) x, cos 2^(L-1) x] -> 3 + 6*multires."""
    outs = [x]
    i = 0
    for k in range(multires):
        f = 2.0 ** k
        for fn in (torch.sin, torch.cos):
            w = 1.0 if ws is None else ws[i]
            outs.append(w * fn(x * f))
            i += 1
    return torch.cat(outs, -1)


# ----------------------------------------------------------------------------------------------
# MLPs on effective weights (model/network.py:89-119, Deformer.py:171-206, RenderNet.py:59-96)
# ----------------------------------------------------------------------------------------------
def weight_norm_effective(g, v):
    """nn.utils.weight_norm default dim=0: W = g * v / ||v||_row."""
    return g * v / v.norm(dim=1, keepdim=True)


def sdf_mlp(x, Ws, bs, pe_w, skip_layer=4, multires=6):
    """Ws[l] [out,in] effective weights.  Returns (sdf [P,1], feat [P,256])."""
    inp = embed(x, multires, pe_w)
    h = inp
    L = len(Ws)
    for l in range(L):
        if l == skip_layer:
            h = torch.cat([h, inp], 1) / np.sqrt(2)
        h = F.linear(h, Ws[l], bs[l])
        if l < L - 1:
            h = F.softplus(h, beta=100)
    return h[:, :1], h[:, 1:]


def translator_mlp(p, cond_rows, Ws, bs, pe_w, multires=6):
    """MLPTranslator: PE(p) ++ cond -> 4x512 ReLU -> 3; returns (p + delta, delta)."""
    h = torch.cat([embed(p, multires, pe_w), cond_rows], 1)
    L = len(Ws)
    for l in range(L):
        h = F.linear(h, Ws[l], bs[l])
        if l < L - 1:
            h = F.relu(h)
    return p + h, h


def render_mlp(p, n, v, feat, Ws, bs, pe_w_v, multires_v=4):
    """RenderingNetwork_view_norm mode 'idr' (multires_n = 0): tanh output."""
    h = torch.cat([p, embed(v, multires_v, pe_w_v), n, feat], -1)
    L = len(Ws)
    for l in range(L):
        h = F.linear(h, Ws[l], bs[l])
        if l < L - 1:
            h = F.relu(h)
    return torch.tanh(h)


# ----------------------------------------------------------------------------------------------
# LBS  (model/Deformer.py:359-445) with bone matrices A [N,24,4,4] as inputs
# ----------------------------------------------------------------------------------------------
def bone_matrices(poses, Js, parents, init_pose):
    """Rodrigues -> kinematic chain -> A = G . init_pose   (Deformer.py:372-405)."""
    N = poses.shape[0]
    R = batch_rodrigues(poses.reshape(-1, 3)).view(N, 24, 3, 3)
    Jb = Js.view(1, 24, 3, 1).expand(N, 24, 3, 1)

    def make_A(Rm, t):
        Rh = F.pad(Rm, [0, 0, 0, 1, 0, 0])
        th = torch.cat([t, torch.ones(Rm.shape[0], 1, 1, dtype=Rm.dtype)], dim=1)
        return torch.cat([Rh, th], 2)

    res = [make_A(R[:, 0], Jb[:, 0])]
    for i in range(1, 24):
        res.append(res[int(parents[i])] @ make_A(R[:, i], Jb[:, i] - Jb[:, int(parents[i])]))
    G = torch.stack(res, 1)
    return G @ init_pose.view(1, 24, 4, 4)


def skin_weights(ws_vox, p, bbox_center, bbox_extend):
    """nps = 2 (p - c)/e -> trilinear sample of the 24-channel voxel (Deformer.py:342-355, 421)."""
    nps = (p - bbox_center.view(1, 3)) / bbox_extend * 2
    w = grid_sample3d_fwd(ws_vox, nps.reshape(1, 1, 1, -1, 3))
    return w.view(ws_vox.shape[1], -1).t()


def lbs_forward(p, A, trans, ws_vox, bbox_center, bbox_extend, batch_inds, tps=None):
    """Canonical -> posed: T = sum_j w_j A_j; out = (T [p;1])_:3 + trans[batch]."""
    w = skin_weights(ws_vox, p if tps is None else tps, bbox_center, bbox_extend)
    T = (w[:, :, None] * A[batch_inds].reshape(-1, 24, 16)).sum(1).view(-1, 4, 4)
    ph = torch.cat([p, torch.ones_like(p[:, :1])], 1)
    return (T @ ph[:, :, None])[:, :3, 0] + trans[batch_inds]


def lbs_inverse(x_obs, A, trans, ws_vox, bbox_center, bbox_extend, batch_inds):
    """North-star 'inverse warp' (SURVEY 8a A5'): weights sampled at the OBSERVATION point,
    [M | t] = sum_j w_j A_j[:3,:],  x_c = M^-1 (x_obs - trans - t) with FastMinv semantics:
    |det M| < 1e-4 -> invalid, x_c = 0."""
    w = skin_weights(ws_vox, x_obs, bbox_center, bbox_extend)
    T = (w[:, :, None] * A[batch_inds].reshape(-1, 24, 16)).sum(1).view(-1, 4, 4)
    Minv, ok = minv3x3_fwd(T[:, :3, :3].contiguous())
    rhs = x_obs - trans[batch_inds] - T[:, :3, 3]
    xc = (Minv @ rhs[:, :, None])[:, :, 0]
    return xc, ok


# ----------------------------------------------------------------------------------------------
# Surface solve step (utils/FindSurfacePs.py:145-207)
# ----------------------------------------------------------------------------------------------
def surface_loss(sdf_val, defp, cam_pos, rays, w1=3.05, w2=1.0):
    direct = defp - cam_pos.view(1, 3)
    up = torch.cross(direct, rays, dim=1)
    loss2 = (up.norm(dim=1) / direct.norm(dim=1)).abs()
    return w1 * sdf_val.abs().view(-1) + w2 * loss2, loss2


def rad2deg_asin(s):
    return torch.arcsin(s) * 180.0 / math.pi


# ----------------------------------------------------------------------------------------------
# interp2x_boundary3d (MCAcc/cuda/interp2x_boundary3d_kernel.cu:9-129 forward, :131-242 backward)
# ----------------------------------------------------------------------------------------------
def interp2x_boundary3d(inp, balance_value):
    """2x-1 upsampling of [N,C,D,H,W] with the reference kernel's per-parity-class formulas and summation order
    (left to right), plus the 'neighbours not all on one side of balance_value' flag."""
    N, C, D, H, W = inp.shape
    out = inp.new_zeros((N, C, 2 * D - 1, 2 * H - 1, 2 * W - 1))
    flag = torch.zeros(out.shape, dtype=torch.bool)

    def sl(odd, n):      # (fine-index slice, [coarse slices of the contributing neighbours along this axis])
        return (slice(1, None, 2), [slice(0, n - 1), slice(1, n)]) if odd else (slice(0, None, 2), [slice(0, n)])

    for oz in (0, 1):
        for oy in (0, 1):
            for ox in (0, 1):
                fz, cz = sl(oz, D)
                fy, cy = sl(oy, H)
                fx, cx = sl(ox, W)
                # neighbour order of the reference kernel: x fastest then y then z, EXCEPT the two 4-neighbour classes
                # with z odd, where z varies fastest (kernel lines 81-112)
                if oz and (ox + oy == 1):
                    order = [(a, b, c) for b in range(len(cy)) for c in range(len(cx)) for a in range(len(cz))]
                else:
                    order = [(a, b, c) for a in range(len(cz)) for b in range(len(cy)) for c in range(len(cx))]
                vs = [inp[:, :, cz[a], cy[b], cx[c]] for a, b, c in order]
                s = vs[0]
                for v in vs[1:]:
                    s = s + v
                n = len(vs)
                out[:, :, fz, fy, fx] = s if n == 1 else (s.double() / float(n)).to(inp.dtype)
                fl = torch.stack([v > balance_value for v in vs])
                flag[:, :, fz, fy, fx] = fl.any(0) & ~fl.all(0)
    return out, flag


def interp2x_boundary3d_backward(grad_out):
    """Adjoint of the upsampling: weights 1, 1/2, 1/4, 1/8 by the number of odd offsets (kernel lines 154-238)."""
    N, C, d, h, w = grad_out.shape
    D, H, W = (d + 1) // 2, (h + 1) // 2, (w + 1) // 2
    gp = torch.nn.functional.pad(grad_out, (1, 1, 1, 1, 1, 1))
    gi = grad_out.new_zeros((N, C, D, H, W))
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                wgt = 0.5 ** (abs(dz) + abs(dy) + abs(dx))
                gi = gi + wgt * gp[:, :, 1 + dz:1 + dz + d:2, 1 + dy:1 + dy + h:2, 1 + dx:1 + dx + w:2]
    return gi


# ----------------------------------------------------------------------------------------------
# implicit-surface gradient, per-ray algebra (engineer/networks/OptimNetwork.py:788-851, 862-873)
# ----------------------------------------------------------------------------------------------
def cross_matrix(v):
    """[v]x with [v]x @ x = v x x (OptimNetwork.py:789-804)."""
    m = torch.zeros((v.shape[0], 3, 3), dtype=v.dtype)
    m[:, 0, 1] = -v[:, 2]; m[:, 0, 2] = v[:, 1]
    m[:, 1, 0] = v[:, 2]; m[:, 1, 2] = -v[:, 0]
    m[:, 2, 0] = -v[:, 1]; m[:, 2, 1] = v[:, 0]
    return m


def surface_grad_coeffs(grad_l_p, grad_f_p, jac, rays, d_minus_c=None):
    """b = [grad_f ; [v]x J], r = grad_l (b^T b)^-1 b^T with the FastMinv rule; returns (-r[:,0], r[:,1:4] (-[v]x),
    r[:,1:4] [d-c]x or None, ok) exactly as the reference composes them."""
    vx = cross_matrix(rays)
    a1 = vx.matmul(jac)
    b = torch.cat([grad_f_p.view(-1, 1, 3), a1], dim=1)
    btb = b.permute(0, 2, 1).matmul(b)
    inv, ok = minv3x3_fwd(btb.contiguous())
    rhs = grad_l_p.view(-1, 1, 3).matmul(inv.matmul(b.permute(0, 2, 1)))      # [N,1,4]
    temp = rhs[:, :, -3:].matmul(-vx).view(-1, 3)
    rg = rhs[:, :, -3:].matmul(cross_matrix(d_minus_c)).view(-1, 3) if d_minus_c is not None else None
    return -rhs[:, 0, 0], temp, rg, ok
