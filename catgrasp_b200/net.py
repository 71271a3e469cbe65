"""Device-resident PointNetCls / PointNetSeg (pointnet2.py:275-329) behind the C ABI."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .weights import pack_blob


class _Net:
    kind = None
    kind_id = None

    def __init__(self, state_dict, device=None):
        self.ctx = _lib.Context.get(device)
        blob, n_out = pack_blob(state_dict, self.kind)
        self.n_out = n_out
        lib = self.ctx.lib
        expect = lib.cg_net_blob_floats(self.kind_id, n_out)
        if expect != blob.size:
            raise _lib.CgError(f"weight blob has {blob.size} floats, library expects {expect}")
        h = C.c_void_p()
        self.ctx.check(lib.cg_net_create(self.ctx.h, self.kind_id, n_out, _lib.ptr(blob), blob.size, C.byref(h)))
        self.h = h
        self.device = torch.device("cuda", self.ctx.device)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.ctx.lib.cg_net_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _dev(self, a, dtype):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(device=self.device, dtype=dtype).contiguous()


class PointNetCls(_Net):
    """forward(x:(B,N,6)) -> logits (B,n_out); mirrors pointnet2.py:289-299 (first return value)."""
    kind, kind_id = "cls", _lib.CG_NET_CLS

    def forward(self, x, return_probs=False):
        x = self._dev(x, torch.float32)
        B, N, D = x.shape
        assert D == 6
        self.ctx.use_torch_stream()
        logits = torch.empty((B, self.n_out), dtype=torch.float32, device=self.device)
        probs = torch.empty_like(logits) if return_probs else None
        self.ctx.check(self.ctx.lib.cg_cls_forward_dev(self.h, _lib.ptr(x), B, N, _lib.ptr(logits), _lib.ptr(probs)))
        return (logits, probs) if return_probs else logits

    __call__ = forward

    def draw_ids_dev(self, M, n_pts, count, seed, first_candidate=0):
        """Counter-based subset draw on the device (cg_draw_ids_dev): (count, n_pts) int32 cuda tensor."""
        self.ctx.use_torch_stream()
        ids = torch.empty((count, n_pts), dtype=torch.int32, device=self.device)
        self.ctx.check(self.ctx.lib.cg_draw_ids_dev(self.ctx.h, int(M), int(n_pts), int(count), C.c_uint64(int(seed)),
                                                    C.c_int64(int(first_candidate)), _lib.ptr(ids)))
        return ids

    def graspq_dev(self, cloud_xyz, cloud_nrm, poses, ids, mean=None, std=None, out=None):
        """Fused transform + forward + softmax on device tensors; returns (probs (B,n_out) f32, label (B,) i32)."""
        M = cloud_xyz.shape[0]
        B = poses.shape[0]
        N = ids.shape[1]
        self.ctx.use_torch_stream()
        if out is not None:
            probs, label = out
            assert probs.is_contiguous() and label.is_contiguous() and poses.is_contiguous() and ids.is_contiguous()
        else:
            probs = torch.empty((B, self.n_out), dtype=torch.float32, device=self.device)
            label = torch.empty((B,), dtype=torch.int32, device=self.device)
        self.ctx.check(self.ctx.lib.cg_graspq_forward_dev(
            self.h, _lib.ptr(cloud_xyz), _lib.ptr(cloud_nrm), M, _lib.ptr(poses), B, _lib.ptr(ids), N,
            _lib.ptr(mean), _lib.ptr(std), _lib.ptr(probs), _lib.ptr(label)))
        return probs, label

    def graspq_host(self, cloud_xyz, cloud_nrm, poses, ids, mean=None, std=None, out_probs=None, out_label=None):
        """Reference-facing blocking call on HOST buffers (numpy or pinned torch CPU tensors)."""
        M = cloud_xyz.shape[0]
        B = poses.shape[0]
        N = ids.shape[1]
        if out_probs is None:
            out_probs = np.empty((B, self.n_out), dtype=np.float32)
        if out_label is None:
            out_label = np.empty((B,), dtype=np.int32)
        self.ctx.use_own_stream()      # blocking host call: never on a (possibly freed) torch stream of an earlier _dev call
        self.ctx.check(self.ctx.lib.cg_graspq_forward_host(
            self.h, _lib.ptr(cloud_xyz), _lib.ptr(cloud_nrm), M, _lib.ptr(poses), B, _lib.ptr(ids), N,
            _lib.ptr(mean), _lib.ptr(std), _lib.ptr(out_probs), _lib.ptr(out_label)))
        return out_probs, out_label


class PointNetSeg(_Net):
    """forward(x:(B,N,6)) -> logits (B,N,n_out); mirrors pointnet2.py:316-329."""
    kind, kind_id = "seg", _lib.CG_NET_SEG

    def forward(self, x):
        x = self._dev(x, torch.float32)
        B, N, D = x.shape
        assert D == 6
        self.ctx.use_torch_stream()
        out = torch.empty((B, N, self.n_out), dtype=torch.float32, device=self.device)
        self.ctx.check(self.ctx.lib.cg_seg_forward_dev(self.h, _lib.ptr(x), B, N, _lib.ptr(out)))
        return out

    __call__ = forward

    def nunocs_host(self, x, bins):
        """x (N,6) float32 host -> (coords (N,3) f32, conf_z (N,) f32, bins (N,3) i32); predicter.py:142-150."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        N = x.shape[0]
        coords = np.empty((N, 3), np.float32)
        conf = np.empty((N,), np.float32)
        b = np.empty((N, 3), np.int32)
        self.ctx.use_own_stream()
        self.ctx.check(self.ctx.lib.cg_nunocs_forward_host(self.h, _lib.ptr(x), N, int(bins), _lib.ptr(coords),
                                                           _lib.ptr(conf), _lib.ptr(b)))
        return coords, conf, b

    def nunocs_dev(self, x, bins):
        x = self._dev(x, torch.float32)
        N = x.shape[0]
        self.ctx.use_torch_stream()
        coords = torch.empty((N, 3), dtype=torch.float32, device=self.device)
        conf = torch.empty((N,), dtype=torch.float32, device=self.device)
        b = torch.empty((N, 3), dtype=torch.int32, device=self.device)
        self.ctx.check(self.ctx.lib.cg_nunocs_forward_dev(self.h, _lib.ptr(x), N, int(bins), _lib.ptr(coords),
                                                          _lib.ptr(conf), _lib.ptr(b)))
        return coords, conf, b
