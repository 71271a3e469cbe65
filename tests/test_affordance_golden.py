"""CPU: the affordance-transfer oracle (oracle/affordance_ref.py) against values produced by the reference's own
compute_grasp_affordance_worker / get_finger_contact_area (tests/golden/make_golden_affordance.py)."""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


def affordance_case():
    """Inputs of the golden case, rebuilt from the seeded synthetic generators (shared with the GPU test)."""
    from catgrasp_b200 import synthetic
    rng = np.random.RandomState(12)
    pts, nrm = synthetic.sample_hex_nut(4000, rng)
    R = synthetic.random_rotation(rng)
    full = pts @ R.T + np.array([0.01, -0.02, 0.70])
    full_n = nrm @ R.T
    affordance = np.clip(0.5 + 0.5 * np.sin(40 * pts[:, 0]) * np.cos(35 * pts[:, 1]), 0, 1)
    vox = np.floor(full / 0.002).astype(np.int64)
    _, inv = np.unique(vox, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    cnt = np.bincount(inv).astype(np.float64)
    down = np.stack([np.bincount(inv, full[:, k]) / cnt for k in range(3)], 1)
    down_n = np.stack([np.bincount(inv, full_n[:, k]) / cnt for k in range(3)], 1)
    boxes = np.array([[0.0, 0.045, -0.010, 0.010]] * 2)
    fmig = np.eye(4)
    fmig[:3, 3] = [-0.01, 0.0, 0.0]
    poses = np.asarray(synthetic.make_candidates(full, full_n, 60, seed=5), np.float64)
    poses[50:] = poses[50:] + np.array([[0, 0, 0, 0.2]] * 3 + [[0, 0, 0, 0]])
    return full, affordance, down, down_n, boxes, fmig, poses


def test_affordance_oracle_matches_reference(golden_dir):
    from oracle import affordance_ref
    g = np.load(os.path.join(golden_dir, "affordance.npz"))
    full, affordance, down, down_n, boxes, fmig, poses = affordance_case()
    p, ncon = affordance_ref.grasp_affordance(poses, fmig, down, down_n, affordance, cKDTree(full), boxes,
                                              [[0, 1, 0], [0, -1, 0]], 0.005)
    assert np.array_equal(np.isnan(p), np.isnan(g["p_T_given_G"])) and np.isnan(p).sum() == 10
    np.testing.assert_allclose(p[~np.isnan(p)], g["p_T_given_G"][~np.isnan(p)], rtol=0, atol=1e-14)
    np.testing.assert_array_equal(ncon, g["n_contacts"])


def test_pointwise_nn_formulation_vs_reference(golden_dir):
    """The kernel's formulation (affordance attached per point once) reproduces the reference's drop pattern and patch
    sizes exactly and its scores up to nearest-neighbour ties (see oracle/affordance_ref.py)."""
    from oracle import affordance_ref
    g = np.load(os.path.join(golden_dir, "affordance.npz"))
    full, affordance, down, down_n, boxes, fmig, poses = affordance_case()
    _, nn = cKDTree(full).query(down)
    p, ncon = affordance_ref.grasp_affordance_pointwise_nn(poses, fmig, down, down_n, affordance[nn], boxes, [1, -1], 0.005)
    assert np.array_equal(np.isnan(p), np.isnan(g["p_T_given_G"]))
    np.testing.assert_array_equal(ncon, g["n_contacts"])
    ok = ~np.isnan(p)
    assert np.abs(p[ok] - g["p_T_given_G"][ok]).max() < 1e-3
    assert (np.abs(p[ok] - g["p_T_given_G"][ok]) == 0).sum() >= 5                 # tie-free patches agree exactly
