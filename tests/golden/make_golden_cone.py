"""Golden vectors for the cone pose enumeration, produced by EXECUTING THE REFERENCE's
dexnet/grasping/grasp_sampler.py::PointConeGraspSampler.sample_grasps (grasp_sampler.py:131-222, :225-298) up to its
call of my_cpp.filterGraspPose, whose ``grasp_poses`` argument is recorded.

Run in the authoring container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_cone.py

Stand-ins (import only unless stated): open3d, trimesh, autolab_core, matplotlib, mpl_toolkits, dexnet.grasping.gripper /
.contacts (meshpy-dependent) are empty stubs; ``my_cpp`` is a recorder (returns no survivors); the gripper is a plain
object with the three attributes the sampler reads.  ``transformations`` (Christoph Gohlke's module, not vendored, not
installed) is the one stand-in that COMPUTES: ``euler_matrix`` is catgrasp_b200.grasp_sampler.euler_matrix, a
restatement of that module's published static-xyz formula -- it is used for six in-plane rotations about x and one
rotation about y (grasp_sampler.py:144,:268).
"""
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")

from catgrasp_b200 import grasp_sampler as mine    # noqa: E402
from catgrasp_b200 import synthetic                # noqa: E402


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


for _m in ["open3d", "trimesh", "autolab_core", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d",
           "dexnet.grasping.gripper", "dexnet.grasping.contacts", "pybullet"]:
    sys.modules[_m] = _Stub(_m)
_tf = types.ModuleType("transformations")
_tf.euler_matrix = mine.euler_matrix
_tf.__all__ = ["euler_matrix"]
sys.modules["transformations"] = _tf


class _Recorder(types.ModuleType):
    def __init__(self):
        super().__init__("my_cpp")
        self.calls = []

    def filterGraspPose(self, grasp_poses, *rest):
        self.calls.append(np.array(grasp_poses))
        return []


_rec = _Recorder()
sys.modules["my_cpp"] = _rec

from dexnet.grasping import grasp_sampler as ref_gs   # noqa: E402  the reference itself


class _Gripper:
    hand_depth = 0.012
    init_bite = 0.002
    trimesh = _Any()
    trimesh_enclosed = _Any()

    def get_grasp_pose_in_gripper_base(self):
        return np.eye(4)


CFG = {"sampling_friction_coef": 0.5, "num_cone_faces": 8, "grasp_samples_per_surface_point": 1, "target_num_grasps": 1,
       "min_contact_dist": 0.0}
CASES = [dict(n_pts=60, seed=4, n_sphere_dir=8, approach_step=0.005, center=False, max_num_samples=9),
         dict(n_pts=40, seed=5, n_sphere_dir=5, approach_step=0.004, center=True, max_num_samples=np.inf),
         # an object of a pile whose flat faces give rank-deficient normal scatter matrices: np.linalg.eig answers some of
         # them with complex pairs and the reference drops those rotations (grasp_sampler.py:273)
         dict(pile=(2400, 6, 43, 3), n_sphere_dir=6, approach_step=0.004, center=False, max_num_samples=12)]


def case_inputs(c):
    if "pile" in c:
        n, k, seed, obj = c["pile"]
        scene = synthetic.make_pile(n, n_objects=k, seed=seed)
        m = scene["object_id"] == obj
        return scene["cloud_xyz"][m].copy(), scene["cloud_normal"][m].copy()
    rng = np.random.RandomState(c["seed"])
    pts, nrm = synthetic.sample_hex_nut(c["n_pts"], rng)
    R = synthetic.random_rotation(rng)
    return pts @ R.T + np.array([0.01, -0.02, 0.70]), nrm @ R.T


def main():
    out = {}
    for k, c in enumerate(CASES):
        pts, nrm = case_inputs(c)
        s = ref_gs.PointConeGraspSampler(_Gripper(), CFG)
        _rec.calls.clear()
        np.random.seed(7)
        s.sample_grasps(background_pts=np.ones((1, 3)) * 99999, points_for_sample=pts.copy(), normals_for_sample=nrm.copy(),
                        num_grasps=np.inf, max_num_samples=c["max_num_samples"], n_sphere_dir=c["n_sphere_dir"],
                        approach_step=c["approach_step"], ee_in_grasp=np.eye(4), cam_in_world=np.eye(4),
                        upper=np.ones(7) * 999, lower=-np.ones(7) * 999, open_gripper_collision_pts=np.ones((1, 3)) * 999999,
                        center_ob_between_gripper=c["center"], filter_ik=False, filter_approach_dir_face_camera=False,
                        adjust_collision_pose=False)
        poses = _rec.calls[0]
        out[f"poses_{k}"] = poses
        out[f"next_rand_{k}"] = np.random.rand(2)
        print("cone case", k, c, "poses", poses.shape)
    sp = ref_gs.hinter_sampling(min_n_pts=1000, radius=1)[0]
    out["hinter_1000"] = sp
    np.savez_compressed(os.path.join(HERE, "cone_poses.npz"), **out)


if __name__ == "__main__":
    main()
