"""CPU ORACLE (test infrastructure only): numpy restatement of run_grasp_simulation.py:50-73
(compute_grasp_affordance_worker) and pybullet_env/env_grasp.py:243-283 (get_finger_contact_area).
PINNED: tests/test_affordance_golden.py checks it against values produced by the reference's own functions
(tests/golden/make_golden_affordance.py)."""
import numpy as np


def _finger_contact_area(box, ob_in_finger, ob_pts, ob_normals, grip_dir, surface_tol):
    """box = (xmin, xmax, zmin, zmax) of the finger mesh vertices.  Returns (surface_pts in the camera frame, dist) or None."""
    grip_dir = np.array(grip_dir, dtype=float)
    grip_dir = grip_dir / np.linalg.norm(grip_dir)
    R, t = ob_in_finger[:3, :3], ob_in_finger[:3, 3]
    cur_pts = (R @ ob_pts.T).T + t
    cur_normals = (R @ ob_normals.T).T
    m = (cur_pts[:, 0] >= box[0]) & (cur_pts[:, 0] <= box[1]) & (cur_pts[:, 2] >= box[2]) & (cur_pts[:, 2] <= box[3])
    if m.sum() == 0:
        return None
    w_pts, w_n = cur_pts[m], cur_normals[m]
    if np.allclose(grip_dir, np.array([0, 1, 0])):
        dist = np.abs(w_pts[:, 1] - w_pts[:, 1].min())
    elif np.allclose(grip_dir, np.array([0, -1, 0])):
        dist = np.abs(w_pts[:, 1] - w_pts[:, 1].max())
    else:
        raise RuntimeError(f"grip_dir={grip_dir}")
    c = dist <= surface_tol
    if c.sum() == 0:
        return None
    dist = dist[c]
    n = w_n[c][np.abs(dist).argmin()].copy()
    n /= np.linalg.norm(n)
    if np.dot(n, grip_dir) > 0:
        return None
    homo = np.concatenate((w_pts[c], np.ones((c.sum(), 1))), axis=-1)
    return (np.linalg.inv(ob_in_finger) @ homo.T).T[:, :3], dist


def grasp_affordance(grasp_poses, finger_mesh_in_grasp, pts_in_cam, normals_in_cam, canonical_affordance, kdtree, finger_boxes,
                     grip_dirs, surface_tol=0.005):
    """p(T|G) per grasp (NaN where the reference returns None) and contact-patch sizes (G, F)."""
    out = np.full(len(grasp_poses), np.nan)
    ncon = np.zeros((len(grasp_poses), len(finger_boxes)), np.int32)
    for i, grasp_in_cam in enumerate(grasp_poses):
        cam_in_finger = np.linalg.inv(finger_mesh_in_grasp) @ np.linalg.inv(grasp_in_cam)
        scores = []
        for f, box in enumerate(finger_boxes):
            r = _finger_contact_area(box, cam_in_finger, pts_in_cam, normals_in_cam, grip_dirs[f], surface_tol)
            if r is None:
                continue
            surface_pts, _ = r
            ncon[i, f] = len(surface_pts)
            _, idx = kdtree.query(surface_pts)
            scores.append(canonical_affordance[idx].mean())
        if scores:
            v = np.array(scores).mean()
            if np.isfinite(v):
                out[i] = v
    return out, ncon


def grasp_affordance_pointwise_nn(grasp_poses, finger_mesh_in_grasp, pts_in_cam, normals_in_cam, aff_of_pts, finger_boxes,
                                  grip_signs, surface_tol=0.005):
    """The formulation the CUDA kernel uses: the nearest-canonical-point affordance is attached to every (down-sampled)
    point once (``aff_of_pts``) instead of being queried per contact patch.  Identical to grasp_affordance() except where a
    point is equidistant from two canonical points (a voxel mean of two points is): the reference's per-patch query breaks
    such ties by the rounding noise of its transform round trip (env_grasp.py:282), so a patch mean can differ by one
    point's affordance / patch size (observed <= 2.3e-4)."""
    T = np.linalg.inv(finger_mesh_in_grasp) @ np.linalg.inv(np.asarray(grasp_poses, np.float64))
    out = np.full(len(T), np.nan)
    ncon = np.zeros((len(T), len(finger_boxes)), np.int32)
    for gi in range(len(T)):
        R, t = T[gi, :3, :3], T[gi, :3, 3]
        q = (R @ pts_in_cam.T).T + t
        tot, nf = 0.0, 0
        for f, sgn in enumerate(grip_signs):
            b = finger_boxes[f]
            m = (q[:, 0] >= b[0]) & (q[:, 0] <= b[1]) & (q[:, 2] >= b[2]) & (q[:, 2] <= b[3])
            if not m.any():
                continue
            y_ext = sgn * (sgn * q[m, 1]).min()
            d = np.abs(q[:, 1] - y_ext)
            idx = np.nonzero(m & (d <= surface_tol))[0]
            n = R @ normals_in_cam[idx[np.argmin(d[idx])]]
            if (n[1] / np.linalg.norm(n)) * sgn > 0:
                continue
            ncon[gi, f] = len(idx)
            tot += aff_of_pts[idx].mean()
            nf += 1
        if nf:
            out[gi] = tot / nf
    return out, ncon
