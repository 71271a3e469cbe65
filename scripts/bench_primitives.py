"""Timing of the per-scene kernels (FPS, ball query, grouping, occupancy grid, RANSAC, filter) on one B200."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from catgrasp_b200 import pointnet2 as pn2, my_cpp
from catgrasp_b200.aligning import estimate9DTransform
from catgrasp_b200.synthetic import make_pile

torch.cuda.set_device(0)


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


out = {}
for N in (20000, 40000):
    sc = make_pile(N, seed=1)
    xyz = torch.from_numpy((sc["cloud_xyz"] - sc["cloud_xyz"].mean(0)).astype(np.float32))[None].cuda()
    start = torch.tensor([0])
    out[f"fps_N{N}_1024_ms"] = timeit(lambda: pn2.farthest_point_sample(xyz, 1024, start_idx=start))
    idx = pn2.farthest_point_sample(xyz, 1024, start_idx=start)
    new_xyz = pn2.index_points(xyz, idx)
    out[f"ball_query_N{N}_S1024_k32_ms"] = timeit(lambda: pn2.query_ball_point(0.004, 32, xyz, new_xyz))
    out[f"sample_and_group_N{N}_ms"] = timeit(lambda: pn2.sample_and_group(1024, 0.004, 32, xyz, xyz, start_idx=start), n=5)
sc = make_pile(20000, seed=1)
K = np.eye(3)
t = time.perf_counter(); occ = my_cpp.makeOccupancyGridFromCloudScan(sc["cloud_xyz"], K, 0.001); out["occupancy_20k_1mm_ms"] = (time.perf_counter() - t) * 1e3
out["occupancy_points"] = int(occ.shape[0])
rng = np.random.RandomState(0)
src = np.round(rng.uniform(-0.5, 0.5, (8192, 3)) / 0.01) * 0.01
T = np.eye(4); T[:3, :3] *= 0.02; T[:3, 3] = [0, 0, 0.7]
tgt = (T @ np.c_[src, np.ones(8192)].T).T[:, :3] + rng.normal(0, 0.0004, (8192, 3))
np.random.seed(0)
t = time.perf_counter(); estimate9DTransform(src, tgt, 0.003, max_iter=10000, max_scale=[0.05] * 3, min_scale=[0.005] * 3, max_dimensions=np.array([1.2] * 3)); out["ransac_10000x8192_total_ms"] = (time.perf_counter() - t) * 1e3
# K5-scale cone enumeration: ~1M poses, centred on a 10k-point object
from catgrasp_b200 import grasp_sampler as gs
rng = np.random.RandomState(0)
S = 1024
surf = rng.uniform(-0.01, 0.01, (S, 3)) + [0, 0, 0.7]
R0s = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(S)])
sph = gs.cone_sphere_points(30)
obj = rng.uniform(-0.01, 0.01, (10000, 3)) + [0, 0, 0.7]
def run():
    return gs.enumerate_poses(surf, R0s, sph, 0.03, 0.005, 0.002, points_for_center=obj)
p64, p32 = run(); torch.cuda.synchronize()
out["cone_poses"] = int(p64.shape[0])
out["cone_enumerate_center_10k_ms"] = timeit(run, n=5)
out["cone_enumerate_only_ms"] = timeit(lambda: gs.enumerate_poses(surf, R0s, sph, 0.03, 0.005, 0.002), n=5)
print(json.dumps(out))
