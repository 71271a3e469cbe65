"""CPU ORACLE (test infrastructure only): numpy restatement of the pose enumeration of
dexnet/grasping/grasp_sampler.py:266-286 (sample_one_surface_point) and :191-203 (center_ob_between_gripper).
PINNED: tests/test_cone_golden.py checks it against poses recorded from the reference's own sample_grasps
(tests/golden/make_golden_cone.py).  The rotation helpers it needs are restated here so that the oracle does not import
the product package."""
import math

import numpy as np


def _rot_x(a):
    si, ci = math.sin(a), math.cos(a)
    return np.array([[1.0, -(ci * 0.0), si * 0.0], [0.0, ci, -si], [-0.0, si, ci]])     # euler_matrix(a,0,0,'sxyz')[:3,:3]


def _normalize_cols(R):
    out = R.copy()
    out /= np.linalg.norm(R, axis=0).reshape(1, 3)
    return out


def _dir_to_rot(direction, ref):
    direction = direction / np.linalg.norm(direction)
    v = np.cross(direction, ref)
    if (v == 0).all():
        return np.eye(3)
    s = np.linalg.norm(v)
    c = direction.dot(ref)
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    if s == 0:
        R = np.array([[1.0, 0, 0], [0, -1, 0], [0, 0, -1]])
    else:
        R = (np.identity(3) + K + K.dot(K) * (1 - c) / (s ** 2)).T
    return _normalize_cols(R)


def enumerate_poses(surface_pts, R0s, sphere_pts, hand_depth, approach_step, init_bite, points_for_center=None):
    poses = []
    for p, R0 in zip(surface_pts, R0s):
        Rs = [R0]
        for sp in sphere_pts:
            R_sphere = _dir_to_rot(sp.copy(), np.array([1, 0, 0]))
            for x_rot in np.arange(0, 180, 30):
                Rs.append(R0 @ R_sphere @ _rot_x(x_rot * np.pi / 180))
        for R in Rs:
            R = _normalize_cols(R)
            a = R[:, 0]
            for d in np.arange(0, hand_depth, approach_step):
                T = np.eye(4)
                T[:3, :3] = R
                T[:3, 3] = p + init_bite * a + a * d
                poses.append(T)
    poses = np.array(poses)
    if points_for_center is not None:
        homo = np.concatenate((points_for_center, np.ones((points_for_center.shape[0], 1))), axis=-1)
        for i in range(len(poses)):
            q = (np.linalg.inv(poses[i]) @ homo.T).T[:, :3]
            c = (q.max(axis=0) + q.min(axis=0)) / 2
            off = np.eye(4)
            off[:3, 3] = [0, c[1], 0]
            poses[i] = poses[i] @ off
    return poses
