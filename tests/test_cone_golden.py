"""CPU: the host half of the cone sampler (catgrasp_b200.grasp_sampler: view sphere, local frames, numpy-RNG stream)
and the enumeration oracle (oracle/cone_ref.py) against poses recorded from the reference's own
PointConeGraspSampler.sample_grasps (tests/golden/make_golden_cone.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from catgrasp_b200 import grasp_sampler as gs  # noqa: E402
from oracle import cone_ref  # noqa: E402

CASES = [dict(n_pts=60, seed=4, n_sphere_dir=8, approach_step=0.005, center=False, max_num_samples=9),
         dict(n_pts=40, seed=5, n_sphere_dir=5, approach_step=0.004, center=True, max_num_samples=np.inf),
         dict(pile=(2400, 6, 43, 3), n_sphere_dir=6, approach_step=0.004, center=False, max_num_samples=12)]
HAND_DEPTH, INIT_BITE = 0.012, 0.002


def case_inputs(c):
    from catgrasp_b200 import synthetic
    if "pile" in c:
        n, k, seed, obj = c["pile"]
        scene = synthetic.make_pile(n, n_objects=k, seed=seed)
        m = scene["object_id"] == obj
        return scene["cloud_xyz"][m].copy(), scene["cloud_normal"][m].copy()
    rng = np.random.RandomState(c["seed"])
    pts, nrm = synthetic.sample_hex_nut(c["n_pts"], rng)
    R = synthetic.random_rotation(rng)
    return pts @ R.T + np.array([0.01, -0.02, 0.70]), nrm @ R.T


def test_view_sphere_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "cone_poses.npz"))
    pts, level = gs.hinter_sampling(1000, radius=1)
    assert pts.shape == g["hinter_1000"].shape
    np.testing.assert_array_equal(pts, g["hinter_1000"])


def test_host_frames_and_enumeration_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "cone_poses.npz"))
    for k, c in enumerate(CASES):
        pts, nrm = case_inputs(c)
        np.random.seed(7)
        ids, R0s, sph = gs.cone_frames(pts, nrm, c["max_num_samples"], c["n_sphere_dir"])
        np.testing.assert_array_equal(np.random.rand(2), g[f"next_rand_{k}"])          # same RNG consumption
        poses = cone_ref.enumerate_poses(pts[ids], R0s, sph, HAND_DEPTH, c["approach_step"], INIT_BITE,
                                         points_for_center=pts if c["center"] else None)
        assert poses.shape == g[f"poses_{k}"].shape
        np.testing.assert_allclose(poses, g[f"poses_{k}"], rtol=0, atol=1e-14)
