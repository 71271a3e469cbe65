"""Developer probe: time the grasp-Q forward (3 trunk launches + FCs) on the K2 shape and, in a CG_BUILD_EXPERIMENTS=1
build, print the per-role cycle counters of the persistent trunk (CG_TRUNK_DEBUG=1) under the timing experiments
CG_TRUNK_EXP (1 = no W3 traffic, 2 = max warps idle, 4 = front warps idle; results are wrong in those runs)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--M", type=int, default=20000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--engine", type=int, default=3)
    args = ap.parse_args()
    from catgrasp_b200.net import PointNetCls
    from catgrasp_b200.synthetic import make_candidates, make_pile, make_state_dict
    torch.cuda.set_device(0)
    net = PointNetCls(make_state_dict("cls", 10, seed=0), device=0)
    net.ctx.set_engine(args.engine)
    scene = make_pile(args.M, seed=0)
    poses = make_candidates(scene["cloud_xyz"], scene["cloud_normal"], args.B, seed=1)
    rng = np.random.RandomState(0)
    ids = rng.randint(0, args.M, (args.B, args.N)).astype(np.int32)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (scene["cloud_xyz"], scene["cloud_normal"], poses, ids)]
    for _ in range(2):
        probs, _ = net.graspq_dev(*d)
    torch.cuda.synchronize()
    net.ctx.profile(True)
    net.ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        probs, _ = net.graspq_dev(*d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    tms, tn = net.ctx.profile_read()
    flop = 2.0 * (139648 + 143744 + 143744) * args.B * args.N
    print(f"exp={os.environ.get('CG_TRUNK_EXP', '0')} engine={args.engine} B={args.B} N={args.N}: {ms:.3f} ms/forward "
          f"({args.B / ms * 1e3:.0f} cand/s), trunks {tms / args.iters:.3f} ms/forward = {flop / (tms / args.iters * 1e-3) / 1e12:.0f} TFLOP/s, "
          f"checksum {float(probs.sum()):.3f}")


if __name__ == "__main__":
    main()
