"""Golden vectors for the HOST half of the path, produced by EXECUTING THE REFERENCE's own
predicter.py / dataset_grasp.py / dataset_nunocs.py / augmentations.py / aligning.py / Utils.py.

Run in the authoring container only (needs /root/reference; the GPU box does not have it):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_hostpath.py

Those modules import packages that are not installed here (open3d, trimesh, transformations, autolab_core, spconv,
matplotlib, the dexnet / PointGroup sub-packages).  None of them is touched by the functions executed below, so
they are replaced by empty stub modules *for the import only*; the arithmetic that runs is the reference's,
unmodified: ``GraspDataset.transform``, ``GraspPredicter.predict_batch``, ``NunocsIsolatedDataset.transform``,
``NormalizeCloud``, ``NunocsPredicter.predict``, ``estimate9DTransform`` (with the real cv2), ``load_model``.
The predicter classes are created with ``__new__`` (their constructors hard-wire ``<code_dir>/artifacts/...`` inside
the read-only reference tree) and given the attributes the constructors would set, the checkpoint going through the
reference's ``load_model``.  ``.cuda()`` is patched to a no-op: the reference forward runs on the CPU in fp32.

Fixtures written (inputs + reference outputs; weights come from catgrasp_b200.synthetic, seeded):
  host_predict_batch.npz   GraspPredicter.predict_batch on two crops (M > n_pts and M < n_pts)
  host_nunocs_random.npz   NunocsPredicter.predict with random weights -> the (None, None) path + captured NOCS cloud
  host_nunocs_lattice.npz  NunocsPredicter.predict success path (lattice weights, see synthetic.make_lattice_seg_state_dict)
  host_ransac9d.npz        estimate9DTransform on noisy correspondences with outliers
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
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return type(name, (_Any,), {})


_ABSENT = ["open3d", "trimesh", "transformations", "autolab_core", "spconv", "spconv.modules", "matplotlib",
           "matplotlib.pyplot", "pybullet",
           "dexnet", "dexnet.grasping", "dexnet.grasping.grasp", "dexnet.grasping.gripper",
           "dexnet.grasping.grasp_sampler",
           "PointGroup", "PointGroup.data", "PointGroup.data.dataset_seg", "PointGroup.model",
           "PointGroup.model.pointgroup", "PointGroup.model.pointgroup.pointgroup", "PointGroup.lib",
           "PointGroup.lib.pointgroup_ops", "PointGroup.lib.pointgroup_ops.functions",
           "PointGroup.lib.pointgroup_ops.functions.pointgroup_ops", "PointGroup.util", "PointGroup.util.config"]
for _m in _ABSENT:
    sys.modules[_m] = _Stub(_m)
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self

import predicter as ref_predicter            # noqa: E402  the reference itself
import aligning as ref_aligning              # noqa: E402
from dataset_grasp import GraspDataset       # noqa: E402
from dataset_nunocs import NunocsIsolatedDataset  # noqa: E402
from pointnet2 import PointNetCls, PointNetSeg    # noqa: E402
from Utils import load_model                 # noqa: E402

from catgrasp_b200 import synthetic          # noqa: E402


LOGIT_GAIN = 12.0


def _f32(a):
    """float64 array holding float32-representable values (keeps the fixtures small and exact)."""
    return np.asarray(a, np.float64).astype(np.float32).astype(np.float64)


def _grasp_predicter(tmp, n_pts, seed):
    d = synthetic.write_artifacts(os.path.join(tmp, f"cls{n_pts}"), "cls", n_pts, seed=seed, logit_gain=LOGIT_GAIN)
    import pickle
    import yaml
    p = ref_predicter.GraspPredicter.__new__(ref_predicter.GraspPredicter)
    with open(f"{d}/config_grasp.yml") as f:
        p.cfg = yaml.safe_load(f)
    with open(f"{d}/normalizer.pkl", "rb") as f:
        tmpn = pickle.load(f)
    p.cfg["mean"], p.cfg["std"] = tmpn["mean"], tmpn["std"]                       # predicter.py:53-58
    ds = GraspDataset.__new__(GraspDataset)
    ds.cfg, ds.phase = p.cfg, "test"
    p.dataset = ds
    p.model = load_model(PointNetCls(n_in=p.cfg["input_channel"], n_out=len(p.cfg["classes"]) - 1),
                         ckpt_dir=f"{d}/best_val.pth.tar").eval()                 # predicter.py:62-64
    return p


def _nunocs_predicter(tmp, sd, mean, std, tag):
    p = ref_predicter.NunocsPredicter.__new__(ref_predicter.NunocsPredicter)
    p.class_name = "nut"
    p.min_scale = [0.005, 0.005, 0.001]                                           # predicter.py:106-108
    p.max_scale = [0.05, 0.05, 0.05]
    p.cfg = {"n_pts": 8192, "input_channel": 6, "ce_loss_bins": 100, "mean": mean, "std": std}
    ds = NunocsIsolatedDataset.__new__(NunocsIsolatedDataset)
    ds.cfg, ds.phase = p.cfg, "test"
    p.dataset = ds
    ck = os.path.join(tmp, f"seg_{tag}.pth.tar")
    torch.save({"epoch": 1, "state_dict": sd}, ck)
    p.model = load_model(PointNetSeg(n_in=6, n_out=300), ckpt_dir=ck).eval()      # predicter.py:128-131
    return p


def golden_predict_batch(tmp):
    out = {}
    rng = np.random.RandomState(11)
    for tag, M in (("big", 1500), ("small", 700)):                                # replace=False / replace=True
        pts, nrm = synthetic.sample_hex_nut(M, rng)
        R = synthetic.random_rotation(rng)
        xyz = _f32(pts @ R.T + np.array([0.01, -0.02, 0.70]))
        nrm = _f32(nrm @ R.T)
        xyz[:5, 2] = 0.05                                                          # dropped by the z >= 0.1 mask (dataset_grasp.py:64)
        poses = []
        for _ in range(12):
            T = np.eye(4)
            T[:3, :3] = synthetic.random_rotation(rng)
            T[:3, 3] = xyz[rng.randint(5, M)] + rng.normal(0, 0.003, 3)
            poses.append(T)
        p = _grasp_predicter(tmp, 1024, seed=21)
        data = {"cloud_xyz": xyz.copy(), "cloud_normal": nrm.copy()}
        np.random.seed(0)
        res = p.predict_batch(data, poses)
        assert np.array_equal(data["cloud_xyz"], xyz), "reference must not mutate the caller's data"
        out[f"{tag}_cloud_xyz"] = xyz.astype(np.float32)
        out[f"{tag}_cloud_normal"] = nrm.astype(np.float32)
        out[f"{tag}_poses"] = np.stack(poses)
        out[f"{tag}_labels"] = np.array([r[0] for r in res], np.int64)
        out[f"{tag}_conf"] = np.array([r[1] for r in res], np.float32)
        out[f"{tag}_probs"] = np.stack([r[2] for r in res]).astype(np.float32)
        out[f"{tag}_next_rand"] = np.random.rand(2)                               # pins the RNG consumption
        assert type(res[0][0]) is np.int64 and res[0][1].dtype == np.float32 and res[0][2].shape == (10,)
    out["artifact_seed"] = np.int64(21)
    out["logit_gain"] = np.float64(LOGIT_GAIN)
    np.savez_compressed(os.path.join(HERE, "host_predict_batch.npz"), **out)
    print("host_predict_batch:", out["big_labels"], out["small_labels"])


class _Recorder:
    """Wraps the reference's estimate9DTransform to record what predict() hands to it (no change to the call)."""

    def __init__(self):
        self.calls = []
        self.real = ref_aligning.estimate9DTransform

    def __call__(self, **kw):
        t, inl = self.real(**kw)
        self.calls.append({"source": kw["source"].copy(), "target": kw["target"].copy(), "thres": kw["PassThreshold"],
                           "transform": None if t is None else t.copy()})
        return t, inl


def golden_nunocs(tmp):
    rng = np.random.RandomState(5)
    mean = np.concatenate([rng.normal(0.5, 0.05, 3), rng.normal(0, 0.05, 3)])
    std = np.concatenate([rng.uniform(0.25, 0.35, 3), rng.uniform(0.5, 0.6, 3)])
    # ---- random weights: NOCS is noise -> no hypothesis survives the gates -> (None, None)
    pts, nrm = synthetic.sample_hex_nut(9000, rng)
    R = synthetic.random_rotation(rng)
    xyz = _f32(pts @ R.T + np.array([0.0, 0.01, 0.71]))
    nrm = _f32(nrm @ R.T)
    xyz[:7, 2] = 0.02
    p = _nunocs_predicter(tmp, synthetic.make_state_dict("seg", 300, seed=31), mean, std, "rand")
    rec = _Recorder()
    ref_predicter.estimate9DTransform = rec
    np.random.seed(0)
    nocs, tf = p.predict({"cloud_xyz": xyz.copy(), "cloud_normal": nrm.copy()})
    assert len(rec.calls) == 2
    src = rec.calls[0]["source"]
    np.savez_compressed(os.path.join(HERE, "host_nunocs_random.npz"), cloud_xyz=xyz.astype(np.float32),
                        cloud_normal=nrm.astype(np.float32), mean=mean, std=std, weight_seed=np.int64(31),
                        keep_ids=p.data_transformed["keep_ids"].astype(np.int32),
                        input=p.data_transformed["input"].astype(np.float32),
                        nocs_bins=np.rint((src + 0.5) * 100).astype(np.uint8),
                        returned_none=np.bool_(nocs is None and tf is None),
                        ransac_none=np.array([c["transform"] is None for c in rec.calls]),
                        next_rand=np.random.rand(2))
    print("host_nunocs_random: returned None =", nocs is None, [c["transform"] is None for c in rec.calls])
    # ---- lattice weights: the success path
    xyz, nrm, g = synthetic.sample_lattice_nut(6000, seed=3)
    xyz[:3, 2] = 0.09                                                              # masked points (z < 0.1)
    p = _nunocs_predicter(tmp, synthetic.make_lattice_seg_state_dict(seed=32, mean=mean, std=std), mean, std, "lat")
    rec = _Recorder()
    ref_predicter.estimate9DTransform = rec
    np.random.seed(0)
    nocs, tf = p.predict({"cloud_xyz": xyz.copy(), "cloud_normal": nrm.copy()})
    assert nocs is not None
    ref_predicter.estimate9DTransform = rec.real
    np.savez_compressed(os.path.join(HERE, "host_nunocs_lattice.npz"), cloud_xyz=xyz, cloud_normal=nrm.astype(np.float32),
                        mean=mean, std=std, weight_seed=np.int64(32),
                        keep_ids=p.data_transformed["keep_ids"].astype(np.int32),
                        nocs_bins=np.rint((nocs + 0.5) * 100).astype(np.uint8), nocs_cloud=nocs.astype(np.float32),
                        transform=tf, nocs_pose=p.nocs_pose, best_ratio=np.float64(p.best_ratio),
                        call_transforms=np.stack([c["transform"] for c in rec.calls]),
                        next_rand=np.random.rand(2))
    print("host_nunocs_lattice: best_ratio", p.best_ratio, "\n", tf)


def golden_ransac(tmp):
    rng = np.random.RandomState(9)
    n = 2048
    src = np.round(rng.uniform(-0.5, 0.5, (n, 3)) / 0.01) * 0.01
    T = np.eye(4)
    T[:3, :3] = synthetic.random_rotation(rng) @ np.diag([0.021, 0.024, 0.009])
    T[:3, 3] = [0.01, -0.02, 0.7]
    tgt = (T @ np.c_[src, np.ones(n)].T).T[:, :3] + rng.normal(0, 0.0008, (n, 3))
    bad = rng.choice(n, n // 5, replace=False)
    src[bad] = np.round(rng.uniform(-0.5, 0.5, (len(bad), 3)) / 0.01) * 0.01
    np.random.seed(3)
    tf, inl = ref_aligning.estimate9DTransform(source=src, target=tgt, PassThreshold=0.003, max_iter=3000,
                                               max_scale=[0.05] * 3, min_scale=[0.005, 0.005, 0.001],
                                               max_dimensions=np.array([1.2] * 3))
    nr = np.random.rand(2)
    # the no-survivor path: scale gates nothing can pass
    np.random.seed(4)
    tf2, inl2 = ref_aligning.estimate9DTransform(source=src, target=tgt, PassThreshold=0.003, max_iter=50,
                                                 max_scale=[0.001] * 3, min_scale=[0.0005] * 3)
    assert tf2 is None and inl2 is None
    np.savez_compressed(os.path.join(HERE, "host_ransac9d.npz"), source=src, target=tgt, transform=tf,
                        inliers=inl.astype(np.int32), next_rand=nr, truth=T)
    print("host_ransac9d: inliers", len(inl), "of", n)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    which = sys.argv[1:] or ["predict_batch", "ransac", "nunocs"]
    with tempfile.TemporaryDirectory() as tmp:
        if "predict_batch" in which:
            golden_predict_batch(tmp)
        if "ransac" in which:
            golden_ransac(tmp)
        if "nunocs" in which:
            golden_nunocs(tmp)


if __name__ == "__main__":
    main()
