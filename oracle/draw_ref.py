"""CPU restatement of catgrasp_b200/csrc/cg_draw.cu (the opt-in device-side subset draw) -- ORACLE, test only.

This is NOT a restatement of reference code: the reference draws with numpy's MT19937 on the host
(/root/reference/dataset_grasp.py:72-73) and a counter-based device draw cannot reproduce those numbers.  The
oracle pins the kernel's integer arithmetic bit for bit; the *distribution* is compared with numpy's own
``np.random.choice`` in tests/test_gpu_parity.py::test_device_draw_statistics.
"""
import numpy as np

_U32 = np.uint32


def _mix32(x):
    x = x.astype(np.uint32)
    x ^= x >> _U32(16)
    x = (x * _U32(0x7feb352d)).astype(np.uint32)
    x ^= x >> _U32(15)
    x = (x * _U32(0x846ca68b)).astype(np.uint32)
    x ^= x >> _U32(16)
    return x


def draw_ids(M, n_pts, count, seed, first_candidate=0):
    """(count, n_pts) int32, same values as cg_draw_ids_dev."""
    with np.errstate(over="ignore"):
        seed_lo, seed_hi = _U32(seed & 0xffffffff), _U32((seed >> 32) & 0xffffffff)
        cand = (np.arange(count, dtype=np.uint64) + np.uint64(first_candidate))
        clo = (cand & np.uint64(0xffffffff)).astype(np.uint32)
        chi = (cand >> np.uint64(32)).astype(np.uint32)
        k0 = _mix32(seed_lo ^ _mix32(clo + _U32(0x9e3779b9)))
        k1 = _mix32(seed_hi ^ _mix32(chi + _U32(0x85ebca6b)) ^ k0)
        n = np.broadcast_to(np.arange(n_pts, dtype=np.uint32)[None, :], (count, n_pts))
        K0 = np.broadcast_to(k0[:, None], (count, n_pts))
        K1 = np.broadcast_to(k1[:, None], (count, n_pts))
        if M < n_pts:
            u = _mix32(K0 ^ _mix32(n * _U32(0x9e3779b1) + K1))
            return ((u.astype(np.uint64) * np.uint64(M)) >> np.uint64(32)).astype(np.int32)
        bits = max(1, int(M - 1).bit_length()) if M > 1 else 1
        h = (bits + 1) >> 1
        hmask = _U32((1 << h) - 1)
        x = n.copy()
        todo = np.ones(x.shape, bool)
        first = True
        while todo.any():
            xs = x[todo]
            L, R = xs >> _U32(h), xs & hmask
            k0s, k1s = K0[todo], K1[todo]
            for r in range(8):
                f = _mix32(R ^ (k1s if r & 1 else k0s) ^ _U32((r * 0x9e3779b9) & 0xffffffff)) & hmask
                L, R = R, L ^ f
            xs = (L << _U32(h)) | R
            x[todo] = xs
            todo[todo] = xs >= _U32(M)
            first = False
        return x.astype(np.int32)
