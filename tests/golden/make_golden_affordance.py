"""Golden vectors for the affordance transfer, produced by EXECUTING THE REFERENCE's
run_grasp_simulation.py::compute_grasp_affordance_worker (:50-73) and pybullet_env/env_grasp.py::get_finger_contact_area
(:243-283) with the real scipy cKDTree.

Run in the authoring container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_affordance.py

Import-only stubs: trimesh, autolab_core, pybullet*, matplotlib, spconv, PointGroup.*, dexnet.grasping.*, my_cpp and the
pybullet_env modules other than env_grasp.  Two stand-ins COMPUTE and are therefore restatements, not the reference:
``open3d`` (absent) is a 20-line numpy point cloud whose ``transform`` applies R p + t to the points and R n to the normals,
which is what Open3D's PointCloud::Transform does and all that Utils.toOpen3dCloud / get_finger_contact_area use of it;
``transformations`` only provides names (nothing on this path calls it).  Grasps are plain objects with the one accessor
the worker calls.
"""
import os
import sys
import types

import numpy as np
from scipy.spatial import cKDTree

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/pybullet_env")


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
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return type(name, (_Any,), {})


for _m in ["trimesh", "autolab_core", "transformations", "pybullet", "pybullet_data", "pybullet_tools", "pybullet_tools.utils",
           "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "spconv", "spconv.modules", "my_cpp",
           "camera", "utils_pybullet", "env_base", "pybullet_env.env_base", "pybullet_env.env", "pybullet_env.utils_pybullet",
           "pybullet_env.env_semantic_grasp", "pybullet_env.camera",
           "dexnet", "dexnet.grasping", "dexnet.grasping.grasp", "dexnet.grasping.gripper", "dexnet.grasping.grasp_sampler",
           "PointGroup", "PointGroup.data", "PointGroup.data.dataset_seg", "PointGroup.model", "PointGroup.model.pointgroup",
           "PointGroup.model.pointgroup.pointgroup", "PointGroup.lib", "PointGroup.lib.pointgroup_ops",
           "PointGroup.lib.pointgroup_ops.functions", "PointGroup.lib.pointgroup_ops.functions.pointgroup_ops",
           "PointGroup.util", "PointGroup.util.config"]:
    sys.modules[_m] = _Stub(_m)


# ---- functional stand-in for the three open3d calls on this path
class _Vec(np.ndarray):
    pass


class _PointCloud:
    def __init__(self):
        self.points = np.zeros((0, 3))
        self.normals = np.zeros((0, 3))
        self.colors = np.zeros((0, 3))

    def transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        self.points = (T[:3, :3] @ np.asarray(self.points).T).T + T[:3, 3]
        if len(self.normals):
            self.normals = (T[:3, :3] @ np.asarray(self.normals).T).T
        return self


_o3d = types.ModuleType("open3d")
_o3d.geometry = types.SimpleNamespace(PointCloud=_PointCloud)
_o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.array(a, dtype=np.float64))
_o3d.io = types.SimpleNamespace()
sys.modules["open3d"] = _o3d

import run_grasp_simulation as ref_run   # noqa: E402  the reference itself (functions only; __main__ guard not entered)

from catgrasp_b200 import synthetic      # noqa: E402


class Grasp:
    def __init__(self, pose):
        self.grasp_pose = pose

    def get_grasp_pose_matrix(self):
        return self.grasp_pose.copy()


class FingerMesh:
    def __init__(self, lo, hi):
        self.vertices = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])], float)


def case():
    rng = np.random.RandomState(12)
    pts, nrm = synthetic.sample_hex_nut(4000, rng)
    R = synthetic.random_rotation(rng)
    full = pts @ R.T + np.array([0.01, -0.02, 0.70])
    full_n = nrm @ R.T
    affordance = np.clip(0.5 + 0.5 * np.sin(40 * pts[:, 0]) * np.cos(35 * pts[:, 1]), 0, 1)      # any per-point score in [0,1]
    vox = np.floor(full / 0.002).astype(np.int64)                                                  # 2 mm voxel means
    _, inv = np.unique(vox, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    cnt = np.bincount(inv).astype(np.float64)
    down = np.stack([np.bincount(inv, full[:, k]) / cnt for k in range(3)], 1)
    down_n = np.stack([np.bincount(inv, full_n[:, k]) / cnt for k in range(3)], 1)
    finger_meshes = [FingerMesh((0.0, -0.004, -0.010), (0.045, 0.004, 0.010)), FingerMesh((0.0, -0.004, -0.010), (0.045, 0.004, 0.010))]
    finger_mesh_in_grasp = np.eye(4)
    finger_mesh_in_grasp[:3, 3] = [-0.01, 0.0, 0.0]
    poses = synthetic.make_candidates(full, full_n, 60, seed=5)
    poses[50:] = poses[50:] + np.array([[0, 0, 0, 0.2]] * 3 + [[0, 0, 0, 0]])                      # far away: no contact
    return full, affordance, down, down_n, finger_meshes, finger_mesh_in_grasp, np.asarray(poses, np.float64)


def main():
    full, affordance, down, down_n, finger_meshes, fmig, poses = case()
    kdtree = cKDTree(full)
    grip_dirs = np.array([[0, 1, 0], [0, -1, 0]])
    finger_ids = np.array([1, 2], dtype=int)
    out = np.full(len(poses), np.nan)
    n_contacts = np.zeros((len(poses), 2), np.int32)
    for i, T in enumerate(poses):
        g = ref_run.compute_grasp_affordance_worker(Grasp(T), fmig, down.copy(), down_n.copy(), affordance, kdtree, grip_dirs,
                                                    finger_meshes, finger_ids)
        if g is not None:
            out[i] = g.p_T_given_G
            for f in g.contacts:
                n_contacts[i, f] = len(g.contacts[f])
    np.savez_compressed(os.path.join(HERE, "affordance.npz"), p_T_given_G=out, n_contacts=n_contacts)
    print("affordance golden: finite", int(np.isfinite(out).sum()), "of", len(out), "contacts", n_contacts.sum(0), out[:8].round(4))


if __name__ == "__main__":
    main()
