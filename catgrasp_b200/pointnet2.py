"""PointNet++ sampling/grouping primitives and the PointNet models, with the
names and call signatures of the reference's ``pointnet2.py`` (file:line cited
per function), executing on B200 through libcatgrasp_b200.so.

All tensors are CUDA tensors; indices are returned as int64 like the reference.
"""
import numpy as np
import torch

from . import _lib
from .net import PointNetCls, PointNetSeg  # noqa: F401  (pointnet2.py:275,302)


def _ctx(t):
    if not t.is_cuda:
        raise _lib.CgError("catgrasp_b200.pointnet2 operates on CUDA tensors only (no CPU fallback)")
    ctx = _lib.Context.get(t.device.index)
    ctx.use_torch_stream()
    return ctx


def _f32(t):
    return t.to(torch.float32).contiguous()


def square_distance(src, dst):
    """pointnet2.py:14-33. src (B,N,3), dst (B,M,3) -> (B,N,M) in the expanded form."""
    src, dst = _f32(src), _f32(dst)
    ctx = _ctx(src)
    B, N, _ = src.shape
    M = dst.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float32, device=src.device)
    ctx.check(ctx.lib.cg_square_distance_dev(ctx.h, _lib.ptr(src), _lib.ptr(dst), B, N, M, _lib.ptr(out)))
    return out


def index_points(points, idx):
    """pointnet2.py:35-51. points (B,N,C), idx (B,S) or (B,S,K) -> (B,S[,K],C)."""
    points = _f32(points)
    ctx = _ctx(points)
    B, N, Cc = points.shape
    shape = list(idx.shape)
    idx32 = idx.reshape(B, -1).to(torch.int32).contiguous()
    S = idx32.shape[1]
    out = torch.empty((B, S, Cc), dtype=torch.float32, device=points.device)
    ctx.check(ctx.lib.cg_index_points_dev(ctx.h, _lib.ptr(points), _lib.ptr(idx32), B, N, Cc, S, _lib.ptr(out)))
    return out.reshape(shape + [Cc])


def farthest_point_sample(xyz, npoint, start_idx=None):
    """pointnet2.py:54-75. xyz (B,N,3) -> (B,npoint) int64.

    ``start_idx`` (B,) makes the reference's ``torch.randint`` start (:66) explicit;
    when None it is drawn with torch.randint exactly like the reference.
    """
    xyz = _f32(xyz)
    ctx = _ctx(xyz)
    B, N, _ = xyz.shape
    if start_idx is None:
        start_idx = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    start = torch.as_tensor(start_idx).to(device=xyz.device, dtype=torch.int32).contiguous()
    out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    ctx.check(ctx.lib.cg_fps_dev(ctx.h, _lib.ptr(xyz), B, N, int(npoint), _lib.ptr(start), _lib.ptr(out)))
    return out.long()


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet2.py:78-98. -> (B,S,nsample) int64, nsample smallest in-ball indices, padded with the first."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    ctx = _ctx(xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty((B, S, nsample), dtype=torch.int32, device=xyz.device)
    r2 = float(np.float32(radius ** 2))   # torch compares the fp32 tensor against float32(radius**2), :93
    ctx.check(ctx.lib.cg_ball_query_dev(ctx.h, r2, int(nsample), _lib.ptr(xyz), _lib.ptr(new_xyz), B, N, S,
                                        _lib.ptr(out)))
    return out.long()


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, start_idx=None):
    """pointnet2.py:101-129."""
    xyz = _f32(xyz)
    ctx = _ctx(xyz)
    B, N, Cc = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint, start_idx=start_idx)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    idx32 = idx.to(torch.int32).contiguous()
    D = 0 if points is None else points.shape[-1]
    pts = None if points is None else _f32(points)
    new_points = torch.empty((B, S, nsample, 3 + D), dtype=torch.float32, device=xyz.device)
    ctx.check(ctx.lib.cg_group_points_dev(ctx.h, _lib.ptr(xyz), _lib.ptr(pts), _lib.ptr(new_xyz), _lib.ptr(idx32),
                                          B, N, D, S, nsample, _lib.ptr(new_points)))
    if returnfps:
        grouped_xyz = index_points(xyz, idx)
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """pointnet2.py:132-149 (pure views/concat, no kernel needed)."""
    B, N, Cc = xyz.shape
    new_xyz = torch.zeros(B, 1, Cc, device=xyz.device)
    grouped_xyz = xyz.view(B, 1, N, Cc)
    if points is not None:
        new_points = torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points
