"""Torch-CPU restatement of the PointNet++ set-abstraction / feature-propagation modules -- ORACLE, test only.

The reference (/root/reference/pointnet2.py) ships the sampling / grouping primitives (:14-149) but no module that
calls them; its model docstrings (:274, :304) cite the upstream PointNet/PointNet++ PyTorch family those primitives
come from.  The two modules below follow that family's published forward passes, built on the primitives:

  PointNetSetAbstraction.forward:  sample_and_group (:101-129) | sample_and_group_all (:132-149) -> permute to
      (B, C, nsample, npoint) -> [Conv2d 1x1 -> BatchNorm2d -> ReLU] x L -> max over nsample
  PointNetFeaturePropagation.forward:  dists = square_distance(xyz1, xyz2) (:14-33); sort; 3 nearest;
      w = (1/(d+1e-8)) / sum; interpolated = sum(index_points(points2, idx) * w) (:35-51) -> cat([points1, interp])
      -> [Conv1d -> BatchNorm1d -> ReLU] x L

PARITY: the primitives are pinned against the reference itself (tests/golden/pn2_primitives.npz).  Module-level
parity is UNPINNED by construction (the reference has no such module to run); tests/golden/pn2_modules.npz was
produced by tests/golden/make_golden_modules.py executing THIS composition with the REFERENCE's own primitive
functions imported from /root/reference/pointnet2.py.
"""
import torch
import torch.nn.functional as F


def _mlp(sd, x, n, conv2d):
    for i in range(n):
        w, b = sd[f"mlp_convs.{i}.weight"].float(), sd[f"mlp_convs.{i}.bias"].float()
        x = F.conv2d(x, w, b) if conv2d else F.conv1d(x, w, b)
        x = F.batch_norm(x, sd[f"mlp_bns.{i}.running_mean"].float(), sd[f"mlp_bns.{i}.running_var"].float(),
                         sd[f"mlp_bns.{i}.weight"].float(), sd[f"mlp_bns.{i}.bias"].float(), training=False, eps=1e-5)
        x = F.relu(x)
    return x


@torch.no_grad()
def set_abstraction(prims, sd, nlayers, npoint, radius, nsample, group_all, xyz, points):
    """prims: module providing sample_and_group / sample_and_group_all (the reference's pointnet2 or a stand-in).
    xyz (B,3,N), points (B,D,N)|None -> new_xyz (B,3,S), new_points (B,C,S), grouped (B,S,K,3+D)."""
    xyz = xyz.permute(0, 2, 1)
    if points is not None:
        points = points.permute(0, 2, 1)
    if group_all:
        new_xyz, new_points = prims.sample_and_group_all(xyz, points)
    else:
        new_xyz, new_points = prims.sample_and_group(npoint, radius, nsample, xyz, points)
    grouped = new_points
    x = new_points.permute(0, 3, 2, 1)                     # (B, C+D, nsample, npoint)
    x = _mlp(sd, x, nlayers, conv2d=True)
    x = torch.max(x, 2)[0]
    return new_xyz.permute(0, 2, 1), x, grouped


@torch.no_grad()
def feature_propagation(prims, sd, nlayers, xyz1, xyz2, points1, points2):
    """xyz1 (B,3,N), xyz2 (B,3,S), points1 (B,D1,N)|None, points2 (B,D2,S) -> (B,C,N), idx (B,N,3), weight (B,N,3)."""
    xyz1 = xyz1.permute(0, 2, 1)
    xyz2 = xyz2.permute(0, 2, 1)
    points2 = points2.permute(0, 2, 1)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx = weight = None
    if S == 1:
        interpolated = points2.repeat(1, N, 1)
    else:
        dists = prims.square_distance(xyz1, xyz2)
        dists, idx = dists.sort(dim=-1)
        dists, idx = dists[:, :, :3], idx[:, :, :3]
        dist_recip = 1.0 / (dists + 1e-8)
        norm = torch.sum(dist_recip, dim=2, keepdim=True)
        weight = dist_recip / norm
        interpolated = torch.sum(prims.index_points(points2, idx) * weight.view(B, N, 3, 1), dim=2)
    if points1 is not None:
        new_points = torch.cat([points1.permute(0, 2, 1), interpolated], dim=-1)
    else:
        new_points = interpolated
    x = _mlp(sd, new_points.permute(0, 2, 1), nlayers, conv2d=False)
    return x, idx, weight
