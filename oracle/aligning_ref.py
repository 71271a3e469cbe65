"""CPU ORACLE (test infrastructure only): 9-DoF RANSAC between the predicted NUNOCS cloud and the observed cloud.

Restates aligning.py:23-119 (``estimateAffine3D``, ``estimate9DTransform_worker``,
``estimate9DTransform``) for the non-kdtree evaluation the predicter uses
(predicter.py:161-164, ``use_kdtree_for_eval=False``).  It consumes the global numpy RNG
exactly like the reference (one ``np.random.choice(len(source), 4, replace=False)`` per
iteration, all drawn up front, aligning.py:91-97) and calls cv2.estimateAffine3D like the reference does.
The product path is catgrasp_b200/aligning.py (CUDA hypothesis scoring); this file is its checker.
PINNED: tests/test_host_golden.py compares it with outputs of the reference's aligning.py run here
(tests/golden/make_golden_hostpath.py -> host_ransac9d.npz, host_nunocs_lattice.npz).
"""
import numpy as np


def _to_homo(pts):
    return np.concatenate((pts, np.ones((pts.shape[0], 1))), axis=-1)


def estimateAffine3D(source, target, PassThreshold):
    """aligning.py:23-33."""
    import cv2
    ret, transform, inliers = cv2.estimateAffine3D(source, target, confidence=0.999, ransacThreshold=PassThreshold)
    tmp = np.eye(4)
    tmp[:3] = transform
    inliers = np.where(inliers > 0)[0]
    return tmp, inliers


def _hypothesis(cur_src, cur_dst, target, PassThreshold, max_scale, min_scale, max_dimensions):
    """aligning.py:36-63: one 4-point affine -> scale gates -> orthogonalised R*diag(scales), or None."""
    transform, _ = estimateAffine3D(source=cur_src, target=cur_dst, PassThreshold=PassThreshold)
    new_transform = transform.copy()
    scales = np.linalg.norm(transform[:3, :3], axis=0)
    if (scales > max_scale).any() or (scales < min_scale).any():
        return None
    R = transform[:3, :3] / scales.reshape(1, 3)
    u, s, vh = np.linalg.svd(R)
    if s.min() < 0.8 or s.max() > 1.2:
        return None
    R = u @ vh
    if np.linalg.det(R) < 0:
        return None
    new_transform[:3, :3] = R @ np.diag(scales)
    transform = new_transform.copy()
    if max_dimensions is not None:
        cloud_at_canonical = (np.linalg.inv(transform) @ _to_homo(target).T).T[:, :3]
        dimensions = cloud_at_canonical.max(axis=0) - cloud_at_canonical.min(axis=0)
        if (dimensions > max_dimensions).any():
            return None
    return transform


def estimate9DTransform(source, target, PassThreshold, max_iter=1000, use_kdtree_for_eval=False,
                        kdtree_eval_resolution=None, max_scale=np.array([99, 99, 99]),
                        min_scale=np.array([0, 0, 0]), max_dimensions=None):
    """aligning.py:83-119.  Returns (best_transform (4,4), inliers) or (None, None)."""
    if use_kdtree_for_eval:
        raise NotImplementedError("kd-tree evaluation (aligning.py:68-79) needs open3d; the predicter never enables it")
    source = np.asarray(source, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    max_scale = np.asarray(max_scale, dtype=np.float64)
    min_scale = np.asarray(min_scale, dtype=np.float64)
    srcs, dsts = [], []
    for _ in range(max_iter):                                   # aligning.py:91-97
        ids = np.random.choice(len(source), size=4, replace=False)
        srcs.append(source[ids])
        dsts.append(target[ids])
    transforms = []
    for i in range(len(srcs)):                                  # aligning.py:99-104
        T = _hypothesis(srcs[i], dsts[i], target, PassThreshold, max_scale, min_scale, max_dimensions)
        if T is not None:
            transforms.append(T)
    if len(transforms) == 0:
        return None, None
    src_h = _to_homo(source)
    ratios = np.empty(len(transforms))
    for i, T in enumerate(transforms):                          # aligning.py:65-67
        errs = np.linalg.norm((T @ src_h.T).T[:, :3] - target, axis=-1)
        ratios[i] = np.sum(errs <= PassThreshold) / len(errs)
    best_id = ratios.argmax()                                   # aligning.py:115
    best_transform = transforms[best_id]
    errs = np.linalg.norm((best_transform @ src_h.T).T[:, :3] - target, axis=-1)
    inliers = np.where(errs <= PassThreshold)[0]
    return best_transform, inliers
