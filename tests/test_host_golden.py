"""CPU: pin the oracle's HOST half (transforms, predicter loops, 9-DoF RANSAC) against vectors produced by executing
the reference's predicter.py / dataset_*.py / augmentations.py / aligning.py (tests/golden/make_golden_hostpath.py)."""
import os

import numpy as np

from catgrasp_b200 import synthetic
from oracle import aligning_ref, transforms_ref


def _cls_cfg_sd(g):
    seed = int(g["artifact_seed"])
    sd = synthetic.make_state_dict("cls", 10, seed=seed, logit_gain=float(g["logit_gain"]))
    rng = np.random.RandomState(seed + 7)          # synthetic.write_artifacts' normalizer
    mean = np.concatenate([rng.normal(0, 0.002, 3), rng.normal(0, 0.05, 3)])
    std = np.concatenate([rng.uniform(0.008, 0.012, 3), rng.uniform(0.5, 0.6, 3)])
    return {"n_pts": 1024, "mean": mean, "std": std}, sd


def test_predict_batch_restatement_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "host_predict_batch.npz"))
    cfg, sd = _cls_cfg_sd(g)
    for tag in ("big", "small"):
        data = {"cloud_xyz": g[f"{tag}_cloud_xyz"].astype(np.float64), "cloud_normal": g[f"{tag}_cloud_normal"].astype(np.float64)}
        np.random.seed(0)
        out = transforms_ref.predict_batch(sd, cfg, data, list(g[f"{tag}_poses"]))
        np.testing.assert_array_equal(np.random.rand(2), g[f"{tag}_next_rand"])        # same RNG consumption
        np.testing.assert_array_equal([o[0] for o in out], g[f"{tag}_labels"])
        np.testing.assert_allclose(np.stack([o[2] for o in out]), g[f"{tag}_probs"], rtol=0, atol=5e-6)
        np.testing.assert_allclose([o[1] for o in out], g[f"{tag}_conf"], rtol=0, atol=5e-6)


def test_nunocs_transform_and_bins_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "host_nunocs_random.npz"))
    cfg = {"n_pts": 8192, "ce_loss_bins": 100, "mean": g["mean"], "std": g["std"]}
    sd = synthetic.make_state_dict("seg", 300, seed=int(g["weight_seed"]))
    data = {"cloud_xyz": g["cloud_xyz"].astype(np.float64), "cloud_normal": g["cloud_normal"].astype(np.float64)}
    np.random.seed(0)
    nocs, conf, logits, dt = transforms_ref.nunocs_predict(sd, cfg, data)
    np.testing.assert_array_equal(dt["keep_ids"], g["keep_ids"])
    np.testing.assert_allclose(dt["input"].astype(np.float32), g["input"], rtol=0, atol=0)
    bins = np.rint((nocs + 0.5) * 100).astype(np.uint8)
    assert (bins != g["nocs_bins"]).mean() < 2e-3          # ties between near-equal logits only


def test_nunocs_lattice_success_path_matches_reference(golden_dir):
    """Lattice weights: identical bins, then the oracle's RANSAC restatement under the same seed returns the
    reference's transforms for both thresholds (predicter.py:160-165)."""
    g = np.load(os.path.join(golden_dir, "host_nunocs_lattice.npz"))
    cfg = {"n_pts": 8192, "ce_loss_bins": 100, "mean": g["mean"], "std": g["std"]}
    sd = synthetic.make_lattice_seg_state_dict(seed=int(g["weight_seed"]), mean=g["mean"], std=g["std"])
    data = {"cloud_xyz": g["cloud_xyz"], "cloud_normal": g["cloud_normal"].astype(np.float64)}
    np.random.seed(0)
    nocs, conf, logits, dt = transforms_ref.nunocs_predict(sd, cfg, data)
    np.testing.assert_array_equal(dt["keep_ids"], g["keep_ids"])
    np.testing.assert_array_equal(np.rint((nocs + 0.5) * 100).astype(np.uint8), g["nocs_bins"])
    np.testing.assert_array_equal(nocs.astype(np.float32), g["nocs_cloud"])
    src = (np.eye(4) @ transforms_ref.to_homo(nocs).T).T[:, :3]
    for i, thres in enumerate([0.003, 0.005]):
        tf, inl = aligning_ref.estimate9DTransform(src, dt["cloud_xyz_original"], thres, max_iter=10000,
                                                   max_scale=[0.05] * 3, min_scale=[0.005, 0.005, 0.001],
                                                   max_dimensions=np.array([1.2] * 3))
        np.testing.assert_allclose(tf, g["call_transforms"][i], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(np.random.rand(2), g["next_rand"])
    np.testing.assert_allclose(g["transform"], g["call_transforms"][0], atol=0)     # first threshold already reaches ratio 1


def test_ransac9d_restatement_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "host_ransac9d.npz"))
    np.random.seed(3)
    tf, inl = aligning_ref.estimate9DTransform(g["source"], g["target"], 0.003, max_iter=3000, max_scale=[0.05] * 3,
                                               min_scale=[0.005, 0.005, 0.001], max_dimensions=np.array([1.2] * 3))
    np.testing.assert_array_equal(np.random.rand(2), g["next_rand"])
    np.testing.assert_allclose(tf, g["transform"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(inl, g["inliers"])
    np.random.seed(4)
    assert aligning_ref.estimate9DTransform(g["source"], g["target"], 0.003, max_iter=50, max_scale=[0.001] * 3,
                                            min_scale=[0.0005] * 3) == (None, None)
