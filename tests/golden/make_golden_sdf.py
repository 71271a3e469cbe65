"""Golden vectors for the SDF lookups and the .sdf file layout, produced by EXECUTING THE REFERENCE's
meshpy/meshpy/sdf.py (Sdf3D._signed_distance, ._signed_distance_batch, .is_any_points_inside: sdf.py:292-359,:377-389)
and meshpy/meshpy/sdf_file.py (SdfFile._read_3d: sdf_file.py:59-87).

Run in the authoring container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sdf.py

open3d / autolab_core / trimesh / transformations are import-only stubs; the ``meshpy`` package object is a bare namespace
pointing at the reference directory so that its ``__init__`` (mesh / rendering imports) is not executed.  ``Sdf3D`` is
created with ``__new__`` and given ``data_``, ``dims_``, ``data_torch`` -- the only attributes the three lookup methods
read; its constructor needs autolab_core's SimilarityTransform, which is why ``transform_pt_obj_to_grid_batch``
(sdf.py:361-373) stays unpinned.  ``SdfFile._read_3d`` ends by constructing an Sdf3D; that call is replaced by a recorder
returning (data, origin, resolution).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, n):
        return _Any()


class _Stub(types.ModuleType):
    __all__ = []
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Any,), {})


for _m in ["open3d", "trimesh", "autolab_core", "transformations"]:
    sys.modules[_m] = _Stub(_m)
_pkg = types.ModuleType("meshpy")
_pkg.__path__ = ["/root/reference/meshpy/meshpy"]
sys.modules["meshpy"] = _pkg

import meshpy.sdf as ref_sdf            # noqa: E402  the reference itself
import meshpy.sdf_file as ref_sdf_file  # noqa: E402

from catgrasp_b200 import synthetic     # noqa: E402
from catgrasp_b200.sdf import write_sdf_file  # noqa: E402


def lookup_coords(dims, seed=0, n=4000):
    rng = np.random.RandomState(seed)
    gc = rng.uniform(-4, dims.max() + 4, (n, 3))
    gc[:50] = np.round(gc[:50])                # exact lattice points
    gc[50:60] = dims - 1                       # the last cell (upper corners out of bounds)
    gc[60:80] += 0.5 - (gc[60:80] % 1)         # exact .5 ties (round-half-even)
    gc[80:90] = -0.25                          # clipped below
    return gc.astype(np.float32).astype(np.float64)


def main():
    g = synthetic.make_gripper_proxy()["open"]
    data = np.ascontiguousarray(g["sdf"], dtype=np.float64)
    dims = np.array(data.shape)
    s = ref_sdf.Sdf3D.__new__(ref_sdf.Sdf3D)
    s.data_ = data
    s.dims_ = dims
    s.data_torch = torch.from_numpy(data).float()
    gc = lookup_coords(dims)
    tri = s._signed_distance(gc.T.copy())                         # the method clips its argument in place
    fast = s._signed_distance(gc.T.copy(), fast=True)
    batch = s._signed_distance_batch(torch.from_numpy(gc.T.copy()).float().unsqueeze(0), fast=True)[0].numpy()
    inside_all = bool(s.is_any_points_inside(gc.T.copy()))
    pos = gc[tri > 0.004]                                          # a set with no point inside
    inside_none = bool(s.is_any_points_inside(pos.T.copy()))
    out = dict(coords=gc.astype(np.float32), trilinear=tri, nearest_clamped=fast, nearest_batch=batch,
               any_inside_all=np.bool_(inside_all), outside_subset=np.nonzero(tri > 0.004)[0].astype(np.int32),
               any_inside_outside_subset=np.bool_(inside_none))
    # ---- file layout
    rng = np.random.RandomState(3)
    small = rng.normal(0, 0.01, (5, 4, 3))
    origin = np.array([-0.01, 0.02, 0.005])
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "t.sdf")
        write_sdf_file(path, small, origin, 0.001)
        ref_sdf_file.sdf.Sdf3D = lambda d, o, r: (d, o, r)         # recorder instead of the autolab-dependent constructor
        rd, ro, rr = ref_sdf_file.SdfFile(path).read()
        out["file_text"] = np.frombuffer(open(path, "rb").read(), np.uint8)
    out.update(file_data=rd, file_origin=ro, file_res=np.float64(rr))
    np.savez_compressed(os.path.join(HERE, "sdf_lookup.npz"), **out)
    print("sdf golden:", tri.shape, "any inside:", inside_all, inside_none, "file dims", rd.shape)


if __name__ == "__main__":
    main()
