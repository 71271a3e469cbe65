"""Float64 numpy restatement of meshpy's Sdf3D lookups (meshpy/meshpy/sdf.py) -- ORACLE.

Used to show that the fp32 predicate of filter_ref.c / the CUDA kernel computes the reference's
formula (sdf.py:292-343 trilinear, :345-359 nearest+clamp, :377-389 any-inside) up to fp32 round-off.
PINNED: tests/test_sdf_golden.py compares these functions with outputs of the reference's own Sdf3D methods executed
under import stubs (tests/golden/make_golden_sdf.py).
"""
import numpy as np

# Sdf3D static corner tables (sdf.py:219-225)
min_coords_x = [0, 2, 3, 5]; max_coords_x = [1, 4, 6, 7]
min_coords_y = [0, 1, 3, 6]; max_coords_y = [2, 4, 5, 7]
min_coords_z = [0, 1, 2, 4]; max_coords_z = [3, 5, 6, 7]


def signed_distance(data, coords):
    """sdf.py:292-343 (fast=False).  data (nx,ny,nz); coords (3,N) grid units."""
    data = np.asarray(data, dtype=np.float64)
    dims = np.array(data.shape)
    coords = np.array(coords, dtype=np.float64).reshape(3, -1)
    for i in range(3):
        coords[i] = np.clip(coords[i], 0, dims[i] - 1)
    min_coords = np.floor(coords)
    max_coords = min_coords + 1
    corners = np.zeros((coords.shape[1], 8, 3), dtype=float)
    corners[:, min_coords_x, 0] = min_coords[0].reshape(-1, 1)
    corners[:, max_coords_x, 0] = max_coords[0].reshape(-1, 1)
    corners[:, min_coords_y, 1] = min_coords[1].reshape(-1, 1)
    corners[:, max_coords_y, 1] = max_coords[1].reshape(-1, 1)
    corners[:, min_coords_z, 2] = min_coords[2].reshape(-1, 1)
    corners[:, max_coords_z, 2] = max_coords[2].reshape(-1, 1)
    sd = np.zeros((coords.shape[1]), dtype=float)
    corners = corners.astype(int)
    for i in range(8):
        cur = corners[:, i]
        oob = (cur < 0).any(axis=1) | (cur >= dims.reshape(1, 3)).any(axis=1)
        inb = ~oob
        vals = np.zeros((len(cur)))
        vals[inb] = data[cur[inb, 0], cur[inb, 1], cur[inb, 2]]
        weights = np.prod(1 - np.abs(cur - coords.T), axis=1)
        sd = sd + weights * vals
    return sd


def signed_distance_nearest(data, coords):
    """sdf.py:345-359 (_signed_distance_batch, fast=True): round, clamp, gather."""
    data = np.asarray(data)
    c = np.round(np.asarray(coords, dtype=np.float64).reshape(3, -1)).astype(int)
    for i in range(3):
        c[i] = np.clip(c[i], 0, data.shape[i] - 1)
    return data[c[0], c[1], c[2]]


def is_any_points_inside(data, coords):
    """sdf.py:377-389: out-of-bounds points are dropped, not clamped."""
    data = np.asarray(data)
    c = np.round(np.asarray(coords, dtype=np.float64).reshape(3, -1)).astype(int)
    keep = np.ones(c.shape[1], bool)
    for i in range(3):
        keep &= (c[i] >= 0) & (c[i] < data.shape[i])
    c = c[:, keep]
    return bool((data[c[0], c[1], c[2]] < 0).any())
