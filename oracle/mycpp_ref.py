"""ctypes wrapper of oracle/_ref/libmycpp_ref.so -- the reference's own my_cpp/common.cpp compiled by oracle/build_ref.py
with the FCL/octomap boundary shimmed (see that file).  ORACLE, test infrastructure only."""
import ctypes as C
import os

import numpy as np

from . import build_ref

_lib = None


def available():
    return build_ref.available() or os.path.exists(build_ref.LIB)


def _load():
    global _lib
    if _lib is None:
        path = build_ref.build()
        if path is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is not present")
        _lib = C.CDLL(path)
        _lib.ref_filterGraspPose.restype = C.c_int
        _lib.ref_makeOccupancyGridFromCloudScan.restype = C.c_int
        _lib.ref_ik_solution_count.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32))
    return a if shape is None else a.reshape(shape)


def filterGraspPose(grasp_poses, symmetry_tfs, nocs_pose, canonical_to_nocs, gripper_in_grasp, filter_dir, adjust,
                    sdf_mode, sdf_open, open_pts, sdf_encl, encl_pts, cam_in_world=None, ee_in_grasp=None,
                    filter_ik=False, upper=None, lower=None, octo_resolution=0.0005, counters=False):
    """The reference's filterGraspPose (common.cpp:156-321): returns the survivors (K,4,4) float32 in the (thread-order
    dependent) order the reference produced them.  ``counters=True`` runs it verbose and also returns the four rejection
    counters it prints (common.cpp:316-319) as a dict."""
    lib = _load()
    if counters:
        import os
        import re
        import tempfile
        lib.ref_set_verbose(C.c_int(1))
        with tempfile.TemporaryFile(mode="w+b") as tf:
            saved = os.dup(1)
            try:
                os.dup2(tf.fileno(), 1)
                out = filterGraspPose(grasp_poses, symmetry_tfs, nocs_pose, canonical_to_nocs, gripper_in_grasp, filter_dir,
                                      adjust, sdf_mode, sdf_open, open_pts, sdf_encl, encl_pts, cam_in_world, ee_in_grasp,
                                      filter_ik, upper, lower, octo_resolution)
            finally:
                os.dup2(saved, 1)
                os.close(saved)
                lib.ref_set_verbose(C.c_int(0))
            tf.seek(0)
            txt = tf.read().decode()
        m = re.search(r"n_approach_dir_rej=(\d+), n_ik_rej=(\d+), n_open_gripper_rej=(\d+), n_close_gripper_rej=(\d+)", txt)
        assert m, f"the reference printed no counters: {txt!r}"
        return out, dict(zip(("approach", "ik", "open", "close"), map(int, m.groups())))
    gp = _f32(grasp_poses, (-1, 16)); st = _f32(symmetry_tfs, (-1, 16))
    p1 = _f32(open_pts, (-1, 3)); p2 = _f32(encl_pts, (-1, 3))
    keep = []
    for slot, (sdf, nv) in enumerate(((sdf_open, 8), (sdf_encl, 16))):      # vertex counts only identify the mesh to the shim
        grid = np.ascontiguousarray(sdf["sdf"], dtype=np.float32)
        dims = np.array(grid.shape, dtype=np.int32)
        org = _f32(sdf["origin"])
        keep += [grid, dims, org]
        lib.ref_register_gripper_sdf(C.c_int(slot), C.c_int(nv), _p(grid), _p(dims), _p(org),
                                     C.c_float(float(np.float32(sdf["res"]))))
    lib.ref_set_sdf_mode(C.c_int(int(sdf_mode)))
    eye = np.eye(4, dtype=np.float32)
    up = np.ascontiguousarray(np.zeros(7) if upper is None else upper, dtype=np.float64)
    lo = np.ascontiguousarray(np.zeros(7) if lower is None else lower, dtype=np.float64)
    cap = gp.shape[0] * st.shape[0]
    out = np.zeros((cap, 4, 4), np.float32)
    n = lib.ref_filterGraspPose(_p(gp), C.c_int(gp.shape[0]), _p(st), C.c_int(st.shape[0]), _p(_f32(nocs_pose, 16)),
                                _p(_f32(canonical_to_nocs, 16)), _p(_f32(eye if cam_in_world is None else cam_in_world, 16)),
                                _p(_f32(eye if ee_in_grasp is None else ee_in_grasp, 16)), _p(_f32(gripper_in_grasp, 16)),
                                C.c_int(int(filter_dir)), C.c_int(int(filter_ik)), C.c_int(int(adjust)), _p(up), _p(lo),
                                C.c_int(up.shape[0]), C.c_int(8), C.c_int(16), _p(p1), C.c_int(p1.shape[0]), _p(p2),
                                C.c_int(p2.shape[0]), C.c_float(octo_resolution), _p(out), C.c_int(cap))
    return out[:n]


def directionVecToRotation(direction, ref):
    lib = _load()
    out = np.zeros(9, np.float32)
    lib.ref_directionVecToRotation(_p(_f32(direction, 3)), _p(_f32(ref, 3)), _p(out))
    return out.reshape(3, 3)


def ik_solution_count(ee_in_base, upper, lower):
    lib = _load()
    up = np.ascontiguousarray(upper, dtype=np.float64); lo = np.ascontiguousarray(lower, dtype=np.float64)
    return lib.ref_ik_solution_count(_p(_f32(ee_in_base, 16)), _p(up), _p(lo), C.c_int(up.shape[0]))


def makeOccupancyGridFromCloudScan(pts, K, resolution, cap=4_000_000):
    lib = _load()
    p = _f32(pts, (-1, 3))
    out = np.zeros((cap, 3), np.float32)
    n = lib.ref_makeOccupancyGridFromCloudScan(_p(p), C.c_int(p.shape[0]), _p(_f32(K, 9)), C.c_float(float(np.float32(resolution))),
                                               _p(out), C.c_int(cap))
    assert n <= cap
    return out[:n]


def sort_poses(poses):
    """Canonical order for comparing survivor sets (the reference's order depends on OpenMP scheduling)."""
    p = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 16)
    order = np.lexsort(p.view(np.uint32).T[::-1])
    return p[order].reshape(-1, 4, 4)
