"""ctypes wrapper of oracle/filter_ref.c -- ORACLE, test infrastructure only."""
import ctypes as C

import numpy as np

from .build_oracle import build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.filter_ref.restype = None
        _lib.sdf_lookup_ref.restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32))
    return a if shape is None else a.reshape(shape)


def filter_ref(grasp_poses, symmetry_tfs, nocs_pose, canonical_to_nocs, gripper_in_grasp, filter_dir, adjust,
               sdf_mode, sdf_open, open_pts, sdf_encl, encl_pts, nthreads=0, margin=0.0, split=False):
    """sdf_* = dict(sdf=(nx,ny,nz) f32, origin=(3,), res=float).  Returns (status u8, offset i8, poses f32 (Q,4,4))."""
    lib = _load()
    gp = _f32(grasp_poses, (-1, 16)); st = _f32(symmetry_tfs, (-1, 16))
    p1 = _f32(open_pts, (-1, 3)); p2 = _f32(encl_pts, (-1, 3))
    G, S = gp.shape[0], st.shape[0]
    Q = G * S
    status = np.zeros(Q, np.uint8); offset = np.zeros(Q, np.int8); poses = np.zeros((Q, 4, 4), np.float32)
    go = np.ascontiguousarray(sdf_open["sdf"], dtype=np.float32)
    do = np.array(go.shape, dtype=np.int32); oo = _f32(sdf_open["origin"])
    if sdf_encl is not None:
        ge = np.ascontiguousarray(sdf_encl["sdf"], dtype=np.float32)
        de = np.array(ge.shape, dtype=np.int32); oe = _f32(sdf_encl["origin"]); re = float(np.float32(sdf_encl["res"]))
    else:
        ge = de = oe = None; re = 1.0
    lib.filter_ref_ms(_p(_f32(nocs_pose, 16)), _p(_f32(canonical_to_nocs, 16)), _p(_f32(gripper_in_grasp, 16)),
                   C.c_int(int(filter_dir)), C.c_int(int(adjust)), C.c_int(int(sdf_mode)), C.c_float(float(np.float32(margin))),
                   C.c_int(int(bool(split))), _p(gp), C.c_int(G), _p(st),
                   C.c_int(S), _p(go), _p(do), _p(oo), C.c_float(float(np.float32(sdf_open["res"]))), _p(p1),
                   C.c_int(p1.shape[0]), _p(ge), _p(de), _p(oe), C.c_float(re), _p(p2), C.c_int(p2.shape[0]),
                   C.c_int(int(nthreads)), _p(status), _p(offset), _p(poses))
    return status, offset, poses


def sdf_lookup_ref(grid, grid_coords, mode):
    lib = _load()
    g = np.ascontiguousarray(grid, dtype=np.float32)
    d = np.array(g.shape, dtype=np.int32)
    gc = np.ascontiguousarray(grid_coords, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(gc.shape[0], np.float32)
    lib.sdf_lookup_ref(_p(g), _p(d), _p(gc), C.c_int(gc.shape[0]), C.c_int(int(mode)), _p(out))
    return out


def occupancy_ref(pts, resolution):
    """oracle/occupancy_ref.c: returns (flags (nx,ny,nz) u8, origin (3,) f32, dims)."""
    lib = _load()
    lib.occupancy_ref.restype = None
    lib.occupancy_geometry_ref.restype = None
    p = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).astype(np.float32)).reshape(-1, 3)
    dims = np.zeros(3, np.int32)
    org = np.zeros(3, np.float32)
    lib.occupancy_geometry_ref(_p(p), C.c_int(p.shape[0]), C.c_float(float(np.float32(resolution))), _p(dims), _p(org))
    flags = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.uint8)
    lib.occupancy_ref(_p(p), C.c_int(p.shape[0]), C.c_float(float(np.float32(resolution))), _p(flags))
    return flags.reshape(dims[0], dims[1], dims[2]), org, dims
