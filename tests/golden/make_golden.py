"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE's pointnet2.py.

Run in the authoring container only (needs /root/reference; the GPU box does not have it):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or golden vectors for this path, so these files are
the pin for oracle/pointnet_ref.py and oracle/pn2_ref.py (and, through them, the CUDA kernels).
Weights come from catgrasp_b200.synthetic.make_state_dict (seeded numpy), so the fixtures only hold
inputs and reference outputs.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")

import pointnet2 as ref  # noqa: E402  the reference itself

from catgrasp_b200.synthetic import make_state_dict  # noqa: E402


def load_ref(model, sd):
    sd = {k.replace("module.", ""): v for k, v in sd.items()}   # Utils.py:141-145
    model.load_state_dict(sd)
    return model.eval()


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(1)
    # ---- PointNetCls
    rng = np.random.RandomState(1)
    x = np.concatenate([rng.normal(0, 1.0, (4, 300, 3)), rng.normal(0, 1.0, (4, 300, 3))], -1).astype(np.float32)
    m = load_ref(ref.PointNetCls(6, 10), make_state_dict("cls", 10, seed=0))
    with torch.no_grad():
        logits, trans_feat = m(torch.from_numpy(x))
    np.savez_compressed(os.path.join(HERE, "pointnet_cls.npz"), x=x, logits=logits.numpy(),
                        probs=logits.softmax(1).numpy(), trans_feat=trans_feat.numpy()[:, :4, :4])
    # ---- PointNetSeg
    x = np.concatenate([rng.uniform(0, 1, (1, 160, 3)), rng.normal(0, 0.6, (1, 160, 3))], -1).astype(np.float32)
    m = load_ref(ref.PointNetSeg(6, 300), make_state_dict("seg", 300, seed=1))
    with torch.no_grad():
        logits, _ = m(torch.from_numpy(x))
    np.savez_compressed(os.path.join(HERE, "pointnet_seg.npz"), x=x, logits=logits.numpy())
    # ---- PN++ primitives
    B, N, S, K = 2, 2048, 128, 16
    xyz = rng.uniform(-0.5, 0.5, (B, N, 3)).astype(np.float32)
    feats = rng.normal(0, 1, (B, N, 3)).astype(np.float32)
    radius = 0.12
    torch.manual_seed(123)
    start = torch.randint(0, N, (B,), dtype=torch.long)          # what pointnet2.py:66 will draw
    torch.manual_seed(123)
    fps = ref.farthest_point_sample(torch.from_numpy(xyz), S)
    assert (fps[:, 0] == start).all()
    new_xyz = ref.index_points(torch.from_numpy(xyz), fps)
    ball = ref.query_ball_point(radius, K, torch.from_numpy(xyz), new_xyz)
    torch.manual_seed(123)
    g_new_xyz, g_new_points, g_grouped_xyz, g_fps = ref.sample_and_group(
        S, radius, K, torch.from_numpy(xyz), torch.from_numpy(feats), returnfps=True)
    sq = ref.square_distance(new_xyz[:, :16], torch.from_numpy(xyz)[:, :256])
    # a camera-frame cloud (z ~ 0.7 m): where the expanded form is noisy
    cam = (rng.uniform(-0.025, 0.025, (1, 1500, 3)) + np.array([0.01, -0.02, 0.7])).astype(np.float32)
    torch.manual_seed(7)
    cam_start = torch.randint(0, 1500, (1,), dtype=torch.long)
    torch.manual_seed(7)
    cam_fps = ref.farthest_point_sample(torch.from_numpy(cam), 64)
    cam_new = ref.index_points(torch.from_numpy(cam), cam_fps)
    cam_ball = ref.query_ball_point(0.004, 8, torch.from_numpy(cam), cam_new)
    cam_sq = ref.square_distance(cam_new, torch.from_numpy(cam))
    # edge cases of Appendix A1/A2 (inclusive boundary, ordered pick, pad, empty ball -> N)
    e_xyz = np.array([[[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [0.5, 0, 0]]], dtype=np.float32)
    e_new = np.array([[[0, 0, 0], [3, 0, 0], [10, 0, 0]]], dtype=np.float32)
    e_ball = ref.query_ball_point(1.0, 4, torch.from_numpy(e_xyz), torch.from_numpy(e_new))
    np.savez_compressed(
        os.path.join(HERE, "pn2_primitives.npz"), xyz=xyz, feats=feats, radius=radius, start=start.numpy(),
        fps=fps.numpy(), ball=ball.numpy(), new_xyz=new_xyz.numpy(), g_new_points=g_new_points.numpy(),
        g_grouped_xyz=g_grouped_xyz.numpy(), sq=sq.numpy(), cam=cam, cam_start=cam_start.numpy(),
        cam_fps=cam_fps.numpy(), cam_ball=cam_ball.numpy(), cam_sq=cam_sq.numpy(), e_xyz=e_xyz, e_new=e_new,
        e_ball=e_ball.numpy())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
