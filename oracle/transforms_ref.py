"""Numpy restatement of the reference's inference-time data transforms and predicter loops -- ORACLE.

Follows dataset_grasp.py:63-91, dataset_nunocs.py:38-65, augmentations.py:70-75,
predicter.py:67-94 and :135-150 line by line (float64 on the host, narrowed to fp32 at the
``.float()`` of predicter.py:84/:142).  PINNED: tests/test_host_golden.py checks this file against vectors made by
running the reference's own predicter / dataset classes (tests/golden/make_golden_hostpath.py).
"""
import copy

import numpy as np
import torch

from .pointnet_ref import pointnet_cls_forward, pointnet_seg_forward


def to_homo(pts):
    return np.concatenate((pts, np.ones((pts.shape[0], 1))), axis=-1)


def grasp_transform(data, grasp_pose, cfg):
    """GraspDataset.transform, phase 'test' (dataset_grasp.py:63-91).  Consumes the global numpy RNG."""
    valid_mask = data["cloud_xyz"][:, 2] >= 0.1
    data["cloud_xyz"] = data["cloud_xyz"][valid_mask].reshape(-1, 3)
    data["cloud_normal"] = data["cloud_normal"][valid_mask].reshape(-1, 3)
    data["cloud_xyz"] = (np.linalg.inv(grasp_pose) @ to_homo(data["cloud_xyz"]).T).T[:, :3]
    data["cloud_normal"] = (np.linalg.inv(grasp_pose[:3, :3]) @ data["cloud_normal"].T).T
    replace = data["cloud_xyz"].shape[0] < cfg["n_pts"]
    ids = np.random.choice(np.arange(data["cloud_xyz"].shape[0]), size=(cfg["n_pts"]), replace=replace)
    data["cloud_xyz"] = data["cloud_xyz"][ids]
    data["cloud_normal"] = data["cloud_normal"][ids].reshape(-1, 3)
    data["cloud_xyz_original"] = copy.deepcopy(data["cloud_xyz"])
    data["input"] = np.concatenate((data["cloud_xyz"], data["cloud_normal"]), axis=-1)
    if "mean" in cfg:
        data["input"] = (data["input"] - cfg["mean"].reshape(1, -1)) / (cfg["std"].reshape(1, -1) + 1e-15)
    data["ids"] = ids
    return data


def predict_batch(sd, cfg, data, grasp_poses, batch_size=200, return_inputs=False):
    """GraspPredicter.predict_batch (predicter.py:67-94) on the CPU."""
    input_datas = []
    for i in range(len(grasp_poses)):
        d = grasp_transform(copy.deepcopy(data), grasp_poses[i], cfg)
        input_datas.append(torch.from_numpy(d["input"]))
    input_datas = torch.stack(input_datas, dim=0)
    n_split = int(np.ceil(len(input_datas) / batch_size))
    ids_split = np.array_split(np.arange(len(input_datas)), n_split)
    out = []
    for ids in ids_split:
        x = input_datas[ids].float()                           # predicter.py:84 (.cuda().float())
        pred = pointnet_cls_forward(sd, x)[0].softmax(dim=1).numpy()
        for b in range(len(pred)):
            cur = pred[b]
            lab = cur.argmax()
            out.append([lab, cur[lab], cur])
    if return_inputs:
        return out, input_datas
    return out


def nunocs_transform(data, cfg):
    """NunocsIsolatedDataset.transform, phase 'test' (dataset_nunocs.py:38-65) + NormalizeCloud (augmentations.py:70-75)."""
    keep_ids = np.arange(data["cloud_xyz"].shape[0])
    valid_mask = data["cloud_xyz"][:, 2] >= 0.1
    keep_ids = keep_ids[valid_mask]
    data["cloud_xyz"] = data["cloud_xyz"][valid_mask]
    replace = data["cloud_xyz"].shape[0] < cfg["n_pts"]
    ids = np.random.choice(np.arange(data["cloud_xyz"].shape[0]), size=(cfg["n_pts"]), replace=replace)
    data["cloud_xyz"] = data["cloud_xyz"][ids]
    keep_ids = keep_ids[ids]
    data["cloud_normal"] = data["cloud_normal"][keep_ids].reshape(-1, 3)
    data["cloud_xyz_original"] = copy.deepcopy(data["cloud_xyz"])
    data["keep_ids"] = keep_ids
    max_xyz = data["cloud_xyz"].max(axis=0)
    min_xyz = data["cloud_xyz"].min(axis=0)
    scale = (max_xyz - min_xyz).max()
    data["cloud_xyz"] = (data["cloud_xyz"] - min_xyz) / (scale + 1e-15)
    data["input"] = np.concatenate((data["cloud_xyz"], data["cloud_normal"]), axis=-1)
    if "mean" in cfg:
        data["input"] = (data["input"] - cfg["mean"].reshape(1, -1)) / (cfg["std"].reshape(1, -1) + 1e-15)
    return data


def nunocs_predict(sd, cfg, data):
    """Network half of NunocsPredicter.predict (predicter.py:136-150)."""
    dt = nunocs_transform(copy.deepcopy(data), cfg)
    x = torch.from_numpy(dt["input"]).float().unsqueeze(0)
    bins = cfg["ce_loss_bins"]
    pred = pointnet_seg_forward(sd, x)[0].reshape(-1, 3, bins)
    bin_resolution = 1 / bins
    pred_coords = pred.argmax(dim=-1).float() * bin_resolution
    probs = pred.softmax(dim=-1)
    confidence_z = torch.gather(probs[:, 2, :], dim=-1, index=pred[:, 2, :].argmax(dim=-1).unsqueeze(-1)).numpy().reshape(-1)
    nocs_cloud = pred_coords.numpy() - 0.5
    return nocs_cloud, confidence_z, pred.numpy(), dt
