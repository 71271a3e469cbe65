"""9-DoF RANSAC between the predicted NUNOCS cloud and the observed cloud (aligning.py:83-119), with all
hypotheses scored on the GPU (csrc/cg_ransac.cu).

The host keeps exactly the reference's RNG consumption -- one ``np.random.choice(len(source), 4, replace=False)`` per
iteration, all drawn up front (aligning.py:91-97) -- and the reference's selection rule (first maximum of the inlier
ratio over the hypotheses that survive the gates, aligning.py:105-117).
"""
import ctypes as C

import numpy as np

from . import _lib


def estimate9DTransform(source, target, PassThreshold, max_iter=1000, use_kdtree_for_eval=False,
                        kdtree_eval_resolution=None, max_scale=np.array([99, 99, 99]),
                        min_scale=np.array([0, 0, 0]), max_dimensions=None):
    """Returns (best_transform (4,4) float64, inliers) or (None, None), like aligning.py:83-119."""
    if use_kdtree_for_eval:
        raise NotImplementedError("kd-tree evaluation (aligning.py:68-79) is never enabled by the predicter")
    source = np.ascontiguousarray(source, dtype=np.float64)
    target = np.ascontiguousarray(target, dtype=np.float64)
    N = source.shape[0]
    ids = np.empty((max_iter, 4), dtype=np.int32)
    for i in range(max_iter):                                   # aligning.py:91-97
        ids[i] = np.random.choice(len(source), size=4, replace=False)
    ctx = _lib.Context.get()
    mins = np.ascontiguousarray(np.asarray(min_scale, dtype=np.float64).reshape(3))
    maxs = np.ascontiguousarray(np.asarray(max_scale, dtype=np.float64).reshape(3))
    mdim = None if max_dimensions is None else np.ascontiguousarray(np.asarray(max_dimensions, dtype=np.float64).reshape(3))
    ratio = np.empty(max_iter, np.float64)
    T = np.empty((max_iter, 4, 4), np.float64)
    valid = np.empty(max_iter, np.uint8)
    ctx.use_own_stream()   # blocking host call
    ctx.check(ctx.lib.cg_ransac9d_host(ctx.h, _lib.ptr(source), _lib.ptr(target), N, _lib.ptr(ids), max_iter,
                                       C.c_double(float(PassThreshold)), _lib.ptr(mins), _lib.ptr(maxs), _lib.ptr(mdim),
                                       _lib.ptr(ratio), _lib.ptr(T), _lib.ptr(valid)))
    keep = np.nonzero(valid)[0]
    if keep.size == 0:
        return None, None
    best = keep[np.argmax(ratio[keep])]                         # first maximum among the survivors (aligning.py:115)
    best_transform = T[best].copy()
    errs = np.linalg.norm((best_transform @ np.c_[source, np.ones(N)].T).T[:, :3] - target, axis=-1)
    inliers = np.where(errs <= PassThreshold)[0]
    return best_transform, inliers
