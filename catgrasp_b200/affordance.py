"""Affordance transfer with the structure of the reference's ``compute_grasp_affordance`` (run_grasp_simulation.py:50-107;
SURVEY.md 8f F4), all grasps of an object in one launch (csrc/cg_affordance.cu).

Host (once per object): the canonical cloud in the camera frame, its 2 mm down-sampled copy with normals, and for every
down-sampled point the affordance of its nearest canonical point (the kd-tree query of run_grasp_simulation.py:62 does not
depend on the grasp: contact points are a subset of the down-sampled cloud).  Device (per grasp): finger-frame transform,
finger-extent test, contact patch, normal test, mean affordance.
"""
import ctypes as C

import numpy as np
import torch
from scipy.spatial import cKDTree

from . import _lib


def finger_boxes_from_meshes(finger_meshes):
    """x / z extent of each finger mesh's vertices (pybullet_env/env_grasp.py:252)."""
    return np.array([[m.vertices[:, 0].min(), m.vertices[:, 0].max(), m.vertices[:, 2].min(), m.vertices[:, 2].max()]
                     for m in finger_meshes], dtype=np.float64)


def compute_grasp_affordance(grasp_poses, finger_mesh_in_grasp, canonical_pts_in_cam, canonical_normals_in_cam,
                             canonical_cloud_in_cam, canonical_affordance, finger_boxes, grip_dirs, surface_tol=0.005,
                             device=0):
    """Returns (p_T_given_G (G,) float64 with NaN where the reference drops the grasp (run_grasp_simulation.py:66-67),
    contact-patch sizes (G, F) int32)."""
    if not torch.cuda.is_available():
        raise _lib.CgError("catgrasp_b200.affordance needs a CUDA device (no CPU fallback)")
    poses = np.asarray(grasp_poses, dtype=np.float64).reshape(-1, 4, 4)
    G = poses.shape[0]
    boxes = np.ascontiguousarray(finger_boxes, dtype=np.float64).reshape(-1, 4)
    F = boxes.shape[0]
    dirs = []
    for d in np.asarray(grip_dirs, dtype=np.float64).reshape(F, 3):
        d = d / np.linalg.norm(d)
        if np.allclose(d, [0, 1, 0]):
            dirs.append(1)
        elif np.allclose(d, [0, -1, 0]):
            dirs.append(-1)
        else:
            raise RuntimeError(f"grip_dir={d}")                            # env_grasp.py:266-267
    if G == 0:
        return np.zeros(0), np.zeros((0, F), np.int32)
    pts = np.ascontiguousarray(canonical_pts_in_cam, dtype=np.float64)
    _, nn = cKDTree(np.asarray(canonical_cloud_in_cam, dtype=np.float64)).query(pts)       # :62, once per object
    aff = np.ascontiguousarray(np.asarray(canonical_affordance, dtype=np.float64)[nn])
    cam_in_finger = np.linalg.inv(np.asarray(finger_mesh_in_grasp, np.float64)) @ np.linalg.inv(poses)   # :52
    ctx = _lib.Context.get(device)
    ctx.use_torch_stream()
    dev = torch.device("cuda", device)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)       # noqa: E731
    d_T, d_pts, d_nrm, d_aff = up(cam_in_finger), up(pts), up(np.asarray(canonical_normals_in_cam, np.float64)), up(aff)
    out_p = torch.empty((G,), dtype=torch.float64, device=dev)
    out_c = torch.zeros((G, 4), dtype=torch.int32, device=dev)
    dirs_c = (C.c_int * F)(*dirs)
    ctx.check(ctx.lib.cg_grasp_affordance_dev(ctx.h, _lib.ptr(d_T), G, _lib.ptr(d_pts), _lib.ptr(d_nrm), _lib.ptr(d_aff),
                                              pts.shape[0], _lib.ptr(boxes), dirs_c, F, C.c_double(float(surface_tol)),
                                              _lib.ptr(out_p), _lib.ptr(out_c)))
    return out_p.cpu().numpy(), out_c.cpu().numpy()[:, :F]
