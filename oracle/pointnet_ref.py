"""Torch-CPU fp32 restatement of the reference's PointNet models (eval mode) -- ORACLE, test only.

Functional form (no nn.Module): takes the checkpoint's state_dict and follows pointnet2.py
operation by operation (conv1d(k=1) -> batch_norm(running stats) -> relu, max over points, FCs),
so results agree with the reference to fp32 round-off.  Cited lines are /root/reference/pointnet2.py.
"""
import torch
import torch.nn.functional as F


def _sd(sd):
    if "state_dict" in sd:
        sd = sd["state_dict"]
    return {k.replace("module.", ""): (v if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in sd.items()}


def _conv_bn(sd, x, conv, bn=None, relu=True):
    y = F.conv1d(x, sd[conv + ".weight"].float(), sd[conv + ".bias"].float())
    if bn is not None:
        y = F.batch_norm(y, sd[bn + ".running_mean"].float(), sd[bn + ".running_var"].float(),
                         sd[bn + ".weight"].float(), sd[bn + ".bias"].float(), training=False, eps=1e-5)
    return F.relu(y) if relu else y


def _fc_bn(sd, x, fc, bn=None, relu=True):
    y = F.linear(x, sd[fc + ".weight"].float(), sd[fc + ".bias"].float())
    if bn is not None:
        y = F.batch_norm(y, sd[bn + ".running_mean"].float(), sd[bn + ".running_var"].float(),
                         sd[bn + ".weight"].float(), sd[bn + ".bias"].float(), training=False, eps=1e-5)
    return F.relu(y) if relu else y


def _stn(sd, x, pre, k):
    """STN3d (:170-185) / STNkd (:208-223): x (B,C,N) -> (B,k,k)."""
    B = x.shape[0]
    x = _conv_bn(sd, x, pre + ".conv1", pre + ".bn1")
    x = _conv_bn(sd, x, pre + ".conv2", pre + ".bn2")
    x = _conv_bn(sd, x, pre + ".conv3", pre + ".bn3")
    x = torch.max(x, 2, keepdim=True)[0].view(-1, 1024)
    x = _fc_bn(sd, x, pre + ".fc1", pre + ".bn4")
    x = _fc_bn(sd, x, pre + ".fc2", pre + ".bn5")
    x = _fc_bn(sd, x, pre + ".fc3", None, relu=False)
    x = x + torch.eye(k, dtype=torch.float32).reshape(1, k * k)
    return x.view(B, k, k)


def encoder(sd, x, global_feat):
    """PointNetEncoder.forward (:241-271). x (B,6,N)."""
    B, D, N = x.shape
    trans = _stn(sd, x, "feat.stn", 3)
    x = x.transpose(2, 1)
    feature = x[:, :, 3:]
    x = torch.bmm(x[:, :, :3], trans)                      # :248
    x = torch.cat([x, feature], dim=2).transpose(2, 1)     # :249-251
    x = _conv_bn(sd, x, "feat.conv1", "feat.bn1")          # :252
    trans_feat = _stn(sd, x, "feat.fstn", 64)              # :255
    x = torch.bmm(x.transpose(2, 1), trans_feat).transpose(2, 1)   # :256-258
    pointfeat = x
    x = _conv_bn(sd, x, "feat.conv2", "feat.bn2")          # :263
    x = _conv_bn(sd, x, "feat.conv3", "feat.bn3", relu=False)   # :264  BN, no ReLU
    x = torch.max(x, 2, keepdim=True)[0].view(-1, 1024)    # :265-266
    if global_feat:
        return x, trans, trans_feat
    x = x.view(-1, 1024, 1).repeat(1, 1, N)                # :270
    return torch.cat([x, pointfeat], 1), trans, trans_feat


@torch.no_grad()
def pointnet_cls_forward(sd, x):
    """PointNetCls.forward (:289-299), eval mode (dropout = identity). x (B,N,6) -> (logits (B,n_out), trans_feat)."""
    sd = _sd(sd)
    x = torch.as_tensor(x, dtype=torch.float32).permute(0, 2, 1)
    g, trans, trans_feat = encoder(sd, x, True)
    y = _fc_bn(sd, g, "fc1", "bn1")
    y = _fc_bn(sd, y, "fc2", "bn2")
    y = _fc_bn(sd, y, "fc3", None, relu=False)
    return y, trans_feat


@torch.no_grad()
def pointnet_seg_forward(sd, x):
    """PointNetSeg.forward (:316-329). x (B,N,6) -> (logits (B,N,n_out), trans_feat)."""
    sd = _sd(sd)
    x = torch.as_tensor(x, dtype=torch.float32).permute(0, 2, 1)
    y, trans, trans_feat = encoder(sd, x, False)
    y = _conv_bn(sd, y, "conv1", "bn1")
    y = _conv_bn(sd, y, "conv2", "bn2")
    y = _conv_bn(sd, y, "conv3", "bn3")
    y = _conv_bn(sd, y, "conv4", None, relu=False)
    return y.permute(0, 2, 1), trans_feat
