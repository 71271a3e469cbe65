"""Numpy fp32 restatement of the reference's PointNet++ primitives (pointnet2.py:14-149) -- ORACLE.

The floating-point forms are the reference's: FPS uses the direct form ((dx^2+dy^2)+dz^2) (:71),
the ball query the expanded form -2<s,d> + |s|^2 + |d|^2 (:30-32).  The dot product of the expanded
form is accumulated as fma(sz,dz, fma(sy,dy, sx*dx)) which is what the BLAS sgemm behind
``torch.matmul`` does for K=3 on an FMA machine; tests/golden pins this against the reference itself.
"""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    # correctly rounded fp32 fma via float64 (exact product of two fp32 fits in 48 bits; one rounding
    # of the fp64 sum then one to fp32 can double-round only in astronomically rare ties)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def sq_expanded(src, dst):
    """square_distance (:14-33): src (S,3), dst (N,3) -> (S,N) fp32."""
    src = src.astype(f32); dst = dst.astype(f32)
    sx, sy, sz = src[:, 0:1], src[:, 1:2], src[:, 2:3]
    dx, dy, dz = dst[None, :, 0], dst[None, :, 1], dst[None, :, 2]
    dot = _fma(np.broadcast_to(sz, (src.shape[0], dst.shape[0])), np.broadcast_to(dz, (src.shape[0], dst.shape[0])),
               _fma(np.broadcast_to(sy, (src.shape[0], dst.shape[0])), np.broadcast_to(dy, (src.shape[0], dst.shape[0])),
                    (sx * dx).astype(f32)))
    ss = ((sx * sx).astype(f32) + (sy * sy).astype(f32)).astype(f32) + (sz * sz).astype(f32)
    dd = ((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32) + (dz * dz).astype(f32)
    out = (f32(-2.0) * dot).astype(f32)
    out = (out + ss).astype(f32)
    out = (out + dd).astype(f32)
    return out


def square_distance(src, dst):
    return np.stack([sq_expanded(src[b], dst[b]) for b in range(src.shape[0])])


def index_points(points, idx):
    """(:35-51)"""
    B = points.shape[0]
    return np.stack([points[b][idx[b]] for b in range(B)])


def farthest_point_sample(xyz, npoint, start_idx):
    """(:54-75) with an explicit start index per cloud."""
    xyz = xyz.astype(f32)
    B, N, _ = xyz.shape
    centroids = np.zeros((B, npoint), dtype=np.int64)
    for b in range(B):
        distance = np.full((N,), 1e10, dtype=f32)
        farthest = int(start_idx[b])
        P = xyz[b]
        for i in range(npoint):
            centroids[b, i] = farthest
            d = P - P[farthest][None]
            dist = ((d[:, 0] * d[:, 0]).astype(f32) + (d[:, 1] * d[:, 1]).astype(f32)).astype(f32) + (d[:, 2] * d[:, 2]).astype(f32)
            mask = dist < distance
            distance[mask] = dist[mask]
            farthest = int(np.argmax(distance))      # first maximum, like torch.max on CPU
    return centroids


def query_ball_point(radius, nsample, xyz, new_xyz):
    """(:78-98): nsample smallest indices with d2 <= r2 (excluded iff d2 > r2), padded with the first; N if empty."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    r2 = f32(radius ** 2)
    out = np.zeros((B, S, nsample), dtype=np.int64)
    for b in range(B):
        d = sq_expanded(new_xyz[b], xyz[b])
        for s in range(S):
            inb = np.nonzero(~(d[s] > r2))[0][:nsample]
            row = np.full((nsample,), N if inb.size == 0 else inb[0], dtype=np.int64)
            row[: inb.size] = inb
            out[b, s] = row
    return out


def sample_and_group(npoint, radius, nsample, xyz, points, start_idx):
    """(:101-129)"""
    fps_idx = farthest_point_sample(xyz, npoint, start_idx)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = (grouped_xyz.astype(f32) - new_xyz.astype(f32)[:, :, None, :]).astype(f32)
    if points is not None:
        new_points = np.concatenate([grouped_xyz_norm, index_points(points, idx)], axis=-1)
    else:
        new_points = grouped_xyz_norm
    return new_xyz, new_points, grouped_xyz, fps_idx
