"""CPU: pin the oracle against vectors produced by executing the reference's pointnet2.py
(tests/golden/make_golden.py)."""
import os

import numpy as np

from catgrasp_b200.synthetic import make_state_dict
from oracle import pn2_ref
from oracle.pointnet_ref import pointnet_cls_forward, pointnet_seg_forward


def test_pointnet_cls_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "pointnet_cls.npz"))
    logits, trans_feat = pointnet_cls_forward(make_state_dict("cls", 10, seed=0), g["x"])
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(logits.softmax(1).numpy(), g["probs"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(trans_feat.numpy()[:, :4, :4], g["trans_feat"], rtol=0, atol=2e-5)


def test_pointnet_seg_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "pointnet_seg.npz"))
    logits, _ = pointnet_seg_forward(make_state_dict("seg", 300, seed=1), g["x"])
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=5e-5)


def test_fps_matches_reference_exactly(golden_dir):
    g = np.load(os.path.join(golden_dir, "pn2_primitives.npz"))
    np.testing.assert_array_equal(pn2_ref.farthest_point_sample(g["xyz"], g["fps"].shape[1], g["start"]), g["fps"])
    np.testing.assert_array_equal(pn2_ref.farthest_point_sample(g["cam"], 64, g["cam_start"]), g["cam_fps"])


def test_ball_query_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "pn2_primitives.npz"))
    ball = pn2_ref.query_ball_point(float(g["radius"]), g["ball"].shape[2], g["xyz"], g["new_xyz"])
    np.testing.assert_array_equal(ball, g["ball"])
    # Appendix A1/A2 edge cases: inclusive boundary, ordered pick, pad with first, empty ball -> N
    e = pn2_ref.query_ball_point(1.0, 4, g["e_xyz"], g["e_new"])
    np.testing.assert_array_equal(e, g["e_ball"])
    assert e[0, 0].tolist() == [0, 1, 4, 0] and e[0, 1].tolist() == [2, 3, 2, 2] and e[0, 2].tolist() == [5, 5, 5, 5]


def test_square_distance_and_camera_frame_band(golden_dir):
    """The expanded form is reproduced to within 1 ulp-scale noise; in camera coordinates (z ~ 0.7 m)
    index parity holds outside a stated band around r^2."""
    g = np.load(os.path.join(golden_dir, "pn2_primitives.npz"))
    sq = pn2_ref.square_distance(g["new_xyz"][:, :16], g["xyz"][:, :256])
    np.testing.assert_allclose(sq, g["sq"], rtol=0, atol=2e-7)
    cam_new = pn2_ref.index_points(g["cam"], g["cam_fps"])
    cam_sq = pn2_ref.square_distance(cam_new, g["cam"])
    assert np.abs(cam_sq - g["cam_sq"]).max() <= 2.4e-7      # <= 2 ulp of |p|^2 ~ 0.49
    ball = pn2_ref.query_ball_point(0.004, 8, g["cam"], cam_new)
    r2 = np.float32(0.004 ** 2)
    band = 2.4e-7
    diff = ball != g["cam_ball"]
    if diff.any():   # every disagreement must involve a point within the band of the radius
        near = np.abs(g["cam_sq"] - r2) <= band
        rows = np.nonzero(diff.any(-1))
        assert all(near[b, s].any() for b, s in zip(*rows))


def test_sample_and_group_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "pn2_primitives.npz"))
    S, K = g["fps"].shape[1], g["ball"].shape[2]
    new_xyz, new_points, grouped_xyz, fps = pn2_ref.sample_and_group(S, float(g["radius"]), K, g["xyz"], g["feats"],
                                                                     g["start"])
    np.testing.assert_array_equal(fps, g["fps"])
    np.testing.assert_array_equal(new_xyz, g["new_xyz"])
    np.testing.assert_array_equal(grouped_xyz, g["g_grouped_xyz"])
    np.testing.assert_array_equal(new_points, g["g_new_points"])
