"""CPU emulation of trunk engine 3 (128->1024 layer as ONE fp16 x fp16 pass, fp32 accumulate):
folded W3 and the post-ReLU X3 activations are rounded to fp16 in all three trunks; everything else fp32.
Reports max |dprob| / |dlogit| against the plain fp32 oracle on the golden-style inputs (decides the engine-3 gate)."""
import sys, os
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from catgrasp_b200.synthetic import make_state_dict
from oracle import pointnet_ref as R

MODE = {"x3": True, "w3": True, "l12": False}

def fold(sd, conv, bn):
    W = sd[conv + ".weight"].double().squeeze(-1); b = sd[conv + ".bias"].double()
    s = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + 1e-5)
    return (W * s[:, None]).float(), ((b - sd[bn + ".running_mean"].double()) * s + sd[bn + ".bias"].double()).float()

def h16(t):
    return t.half().float()

orig = R._conv_bn
def conv_bn_emul(sd, x, conv, bn=None, relu=True):
    if conv.endswith("conv3") and bn is not None and x.shape[1] == 128:
        W, b = fold(sd, conv, bn)
        xx = h16(x.clamp(max=65504.)) if MODE["x3"] else x
        WW = h16(W) if MODE["w3"] else W
        y = torch.einsum("oc,bcn->bon", WW.double(), xx.double()).float() + b[None, :, None]
        return F.relu(y) if relu else y
    return orig(sd, x, conv, bn, relu)

def run(kind, n_out, seed, x, gain=1.0):
    sd = R._sd(make_state_dict(kind, n_out, seed=seed))
    f = R.pointnet_cls_forward if kind == "cls" else R.pointnet_seg_forward
    R._conv_bn = orig
    y0 = f(sd, x)[0] * gain
    R._conv_bn = conv_bn_emul
    y1 = f(sd, x)[0] * gain
    R._conv_bn = orig
    return y0, y1

rng = np.random.RandomState(1)
x = np.concatenate([rng.normal(0, 1.0, (64, 1024, 3)), rng.normal(0, 1.0, (64, 1024, 3))], -1).astype(np.float32)
for gain in (1.0, 12.0):
    y0, y1 = run("cls", 10, 0, x, gain)
    print(f"cls gain {gain}: max|dlogit| {float((y0-y1).abs().max()):.3e}  max|dprob| {float((y0.softmax(1)-y1.softmax(1)).abs().max()):.3e}")
xs = np.concatenate([rng.uniform(0, 1, (1, 8192, 3)), rng.normal(0, 0.6, (1, 8192, 3))], -1).astype(np.float32)
y0, y1 = run("seg", 300, 1, xs)
d = (y0 - y1).abs()
b0 = y0.reshape(-1, 3, 100).argmax(-1); b1 = y1.reshape(-1, 3, 100).argmax(-1)
print(f"seg: max|dlogit| {float(d.max()):.3e}  bins flipped {int((b0!=b1).sum())} of {b0.numel()}")
