"""CPU: the SDF oracles (oracle/sdf_ref.py float64, oracle/filter_ref.c float32) and the .sdf parser against outputs of the
reference's own meshpy/meshpy/sdf.py and sdf_file.py (tests/golden/make_golden_sdf.py)."""
import os

import numpy as np

from catgrasp_b200.sdf import parse_sdf_file
from catgrasp_b200.synthetic import make_gripper_proxy
from oracle import filter_ref, sdf_ref


def test_sdf_lookups_match_reference(golden_dir):
    g_ = np.load(os.path.join(golden_dir, "sdf_lookup.npz"))
    grid = make_gripper_proxy()["open"]["sdf"]
    gc = g_["coords"].astype(np.float64)
    np.testing.assert_array_equal(sdf_ref.signed_distance(grid, gc.T), g_["trilinear"])        # same float64 formula
    np.testing.assert_array_equal(sdf_ref.signed_distance_nearest(grid, gc.T), g_["nearest_clamped"])
    np.testing.assert_allclose(sdf_ref.signed_distance_nearest(grid, gc.T), g_["nearest_batch"], rtol=0, atol=1e-7)
    assert bool(sdf_ref.is_any_points_inside(grid, gc.T)) == bool(g_["any_inside_all"])
    assert bool(sdf_ref.is_any_points_inside(grid, gc[g_["outside_subset"]].T)) == bool(g_["any_inside_outside_subset"])
    # the fp32 oracle shared with the CUDA kernel computes the same values up to fp32 round-off
    tri32 = filter_ref.sdf_lookup_ref(grid, gc.astype(np.float32), 0)
    near32 = filter_ref.sdf_lookup_ref(grid, gc.astype(np.float32), 1)
    assert np.abs(tri32 - g_["trilinear"]).max() < 1e-6
    np.testing.assert_allclose(near32, g_["nearest_clamped"], rtol=0, atol=1e-7)


def test_sdf_file_parser_matches_reference_reader(golden_dir, tmp_path):
    g_ = np.load(os.path.join(golden_dir, "sdf_lookup.npz"))
    path = tmp_path / "t.sdf"
    path.write_bytes(g_["file_text"].tobytes())
    data, origin, res = parse_sdf_file(str(path))
    np.testing.assert_array_equal(data, g_["file_data"])          # data[i][j][k], i fastest in the file (sdf_file.py:78-83)
    np.testing.assert_array_equal(origin, g_["file_origin"])
    assert res == float(g_["file_res"])
