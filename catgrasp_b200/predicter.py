"""GraspPredicter / NunocsPredicter with the reference's call surface (predicter.py:39-203).

Host side (numpy, identical RNG consumption to the reference):
  * z >= 0.1 mask and the per-candidate ``np.random.choice`` subset (dataset_grasp.py:64,72-73;
    dataset_nunocs.py:40-44) -- only the *indices* are drawn on the host;
  * NUNOCS min/max-extent normalisation (augmentations.py:70-75);
  * the 9-DoF RANSAC (aligning.py:83-119) stays a host stage.
Device side (libcatgrasp_b200.so): per-candidate rigid transform + normalisation + PointNet forward
+ softmax / argmax post-processing.
"""
import copy
import os
import pickle

import numpy as np
import yaml

from .net import PointNetCls, PointNetSeg
from .weights import load_checkpoint

_CODE_DIR = os.environ.get("CATGRASP_CODE_DIR", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def to_homo(pts):
    """Utils.py:396-402."""
    assert len(pts.shape) == 2, f"pts.shape: {pts.shape}"
    return np.concatenate((pts, np.ones((pts.shape[0], 1))), axis=-1)


def _load_artifacts(artifact_dir, cfg_name):
    with open(f"{artifact_dir}/{cfg_name}", "r") as ff:
        cfg = yaml.safe_load(ff)
    normalizer_dir = f"{artifact_dir}/normalizer.pkl"
    if os.path.exists(normalizer_dir):          # predicter.py:53-58 / :122-126
        with open(normalizer_dir, "rb") as ff:
            tmp = pickle.load(ff)
        cfg["mean"] = np.asarray(tmp["mean"], dtype=np.float64)
        cfg["std"] = np.asarray(tmp["std"], dtype=np.float64)
    return cfg


def draw_subsample_ids(M, n_pts, count=None):
    """The reference's per-sample index draw (dataset_grasp.py:72-73, dataset_nunocs.py:43-44):
    ``np.random.choice(np.arange(M), size=n_pts, replace=M < n_pts)`` from the GLOBAL numpy RNG,
    once per candidate, in candidate order."""
    replace = M < n_pts
    pop = np.arange(M)
    if count is None:
        return np.random.choice(pop, size=(n_pts), replace=replace).astype(np.int32)
    out = np.empty((count, n_pts), dtype=np.int32)
    for i in range(count):
        out[i] = np.random.choice(pop, size=(n_pts), replace=replace)
    return out


class GraspPredicter:
    """predicter.py:39-94."""

    class_name_to_artifact_id = {"nut": 47, "hnm": 51, "screw": 50}

    def __init__(self, class_name, artifact_dir=None, device=None):
        artifact_id = self.class_name_to_artifact_id[class_name]
        if artifact_dir is None:
            artifact_dir = f"{_CODE_DIR}/artifacts/artifacts-{artifact_id}"
        print("GraspPredicter artifact_dir", artifact_dir)
        self.class_name = class_name
        self.cfg = _load_artifacts(artifact_dir, "config_grasp.yml")
        n_out = len(self.cfg["classes"]) - 1
        assert self.cfg["input_channel"] == 6, "the B200 path implements the shipped 6-channel input"
        sd = load_checkpoint(f"{artifact_dir}/best_val.pth.tar")
        print("Load ckpt from {}/best_val.pth.tar".format(artifact_dir))
        self.model = PointNetCls(sd, device=device)
        assert self.model.n_out == n_out, f"checkpoint has {self.model.n_out} classes, config says {n_out}"

    def predict_batch(self, data, grasp_poses, ids=None):
        """predicter.py:67-94.  Returns list of [label np.int64, confidence np.float32, probs (n_out,) f32].

        ``data`` is not modified (the reference deep-copies it, :72).  ``ids`` (B,n_pts) overrides the
        numpy-RNG draw for reproducible comparisons.
        """
        B = len(grasp_poses)
        if B == 0:
            return []
        xyz = np.asarray(data["cloud_xyz"], dtype=np.float64)
        nrm = np.asarray(data["cloud_normal"], dtype=np.float64)
        valid_mask = xyz[:, 2] >= 0.1                                   # dataset_grasp.py:64
        xyz = np.ascontiguousarray(xyz[valid_mask].reshape(-1, 3))
        nrm = np.ascontiguousarray(nrm[valid_mask].reshape(-1, 3))
        n_pts = int(self.cfg["n_pts"])
        if ids is None:
            ids = draw_subsample_ids(xyz.shape[0], n_pts, count=B)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        poses = np.ascontiguousarray(np.asarray(grasp_poses, dtype=np.float64).reshape(B, 4, 4))
        mean = np.ascontiguousarray(self.cfg["mean"].reshape(-1)) if "mean" in self.cfg else None
        std = np.ascontiguousarray(self.cfg["std"].reshape(-1)) if "std" in self.cfg else None
        probs, _ = self.model.graspq_host(xyz, nrm, poses, ids, mean, std)
        out = []
        for b in range(B):                                              # predicter.py:87-91
            cur_pred = probs[b]
            pred_label = cur_pred.argmax()
            out.append([pred_label, cur_pred[pred_label], cur_pred])
        return out


class NunocsPredicter:
    """predicter.py:98-203."""

    class_name_to_artifact_id = {"nut": 78, "hnm": 73, "screw": 76}

    def __init__(self, class_name, artifact_dir=None, device=None):
        self.class_name = class_name
        if class_name == "nut":                                         # predicter.py:106-114
            self.min_scale = [0.005, 0.005, 0.001]
            self.max_scale = [0.05, 0.05, 0.05]
        else:
            self.min_scale = [0.005, 0.005, 0.005]
            self.max_scale = [0.15, 0.05, 0.05]
        artifact_id = self.class_name_to_artifact_id[class_name]
        if artifact_dir is None:
            artifact_dir = f"{_CODE_DIR}/artifacts/artifacts-{artifact_id}"
        print("NunocsPredicter artifact_dir", artifact_dir)
        self.cfg = _load_artifacts(artifact_dir, "config_nunocs.yml")
        sd = load_checkpoint(f"{artifact_dir}/best_val.pth.tar")
        self.model = PointNetSeg(sd, device=device)
        assert self.model.n_out == 3 * self.cfg["ce_loss_bins"]
        self.ransac_max_iter = 10000

    def transform(self, data, ids=None):
        """NunocsIsolatedDataset.transform in 'test' phase (dataset_nunocs.py:38-65)."""
        keep_ids = np.arange(data["cloud_xyz"].shape[0])
        valid_mask = data["cloud_xyz"][:, 2] >= 0.1
        keep_ids = keep_ids[valid_mask]
        data["cloud_xyz"] = data["cloud_xyz"][valid_mask]
        if ids is None:
            ids = draw_subsample_ids(data["cloud_xyz"].shape[0], int(self.cfg["n_pts"]))
        data["cloud_xyz"] = data["cloud_xyz"][ids]
        keep_ids = keep_ids[ids]
        data["cloud_nocs"] = data["cloud_nocs"][keep_ids].reshape(-1, 3) / 255.0
        data["cloud_rgb"] = data["cloud_rgb"][keep_ids].reshape(-1, 3)
        data["cloud_normal"] = data["cloud_normal"][keep_ids].reshape(-1, 3)
        data["cloud_xyz_original"] = copy.deepcopy(data["cloud_xyz"])
        data["keep_ids"] = keep_ids
        max_xyz = data["cloud_xyz"].max(axis=0)                         # augmentations.py:70-75
        min_xyz = data["cloud_xyz"].min(axis=0)
        scale = (max_xyz - min_xyz).max()
        data["cloud_xyz"] = (data["cloud_xyz"] - min_xyz) / (scale + 1e-15)
        data["input"] = np.concatenate((data["cloud_xyz"], data["cloud_normal"]), axis=-1)
        if "mean" in self.cfg:
            data["input"] = (data["input"] - self.cfg["mean"].reshape(1, -1)) / (self.cfg["std"].reshape(1, -1) + 1e-15)
        if "color_file" in data:
            del data["color_file"]
        return data

    def predict_nocs(self, data, ids=None):
        """Network half of predict (predicter.py:136-150): returns (nocs_cloud (N,3) f32, confidence_z (N,))."""
        data["cloud_nocs"] = np.zeros(data["cloud_xyz"].shape)
        data["cloud_rgb"] = np.zeros(data["cloud_xyz"].shape)
        data_transformed = self.transform(copy.deepcopy(data), ids=ids)
        self.data_transformed = data_transformed
        x = np.ascontiguousarray(data_transformed["input"], dtype=np.float64).astype(np.float32)
        coords, conf_z, bins = self.model.nunocs_host(x, int(self.cfg["ce_loss_bins"]))
        self.confidence_z = conf_z
        self.pred_bins = bins
        return coords, conf_z

    def predict(self, data, ids=None):
        """predicter.py:135-203: (nocs_cloud, transform) or (None, None)."""
        from .aligning import estimate9DTransform
        nocs_cloud, _ = self.predict_nocs(data, ids=ids)
        ori_cloud = self.data_transformed["cloud_xyz_original"]
        nocs_cloud_down = copy.deepcopy(nocs_cloud)
        ori_cloud_down = copy.deepcopy(ori_cloud)
        best_ratio = 0
        best_transform = None
        best_symmetry_tf = None
        for symmetry_tf in [np.eye(4)]:
            tmp_nocs_cloud_down = (symmetry_tf @ to_homo(nocs_cloud_down).T).T[:, :3]
            for thres in [0.003, 0.005]:
                transform, inliers = estimate9DTransform(
                    source=tmp_nocs_cloud_down, target=ori_cloud_down, PassThreshold=thres,
                    max_iter=self.ransac_max_iter, max_scale=self.max_scale, min_scale=self.min_scale,
                    max_dimensions=np.array([1.2, 1.2, 1.2]))
                if transform is None:
                    continue
                if np.linalg.det(transform[:3, :3]) < 0:
                    continue
                transformed = (transform @ to_homo(tmp_nocs_cloud_down).T).T[:, :3]
                err_thres = 0.003
                errs = np.linalg.norm(transformed - ori_cloud_down, axis=1)
                ratio = np.sum(errs <= err_thres) / len(errs)
                if ratio > best_ratio:
                    best_ratio = ratio
                    best_symmetry_tf = symmetry_tf
                    best_transform = transform.copy()
        if best_transform is None:
            return None, None
        self.best_ratio = best_ratio
        transform = best_transform
        self.nocs_pose = transform.copy()
        nocs_cloud = (best_symmetry_tf @ to_homo(nocs_cloud).T).T[:, :3]
        return nocs_cloud, transform
