"""CPU: how often does the gripper-SDF predicate (product, oracle/filter_ref.c) agree with the mesh-vs-voxel predicate the
reference gets from FCL + octomap?  The latter is only available as a restatement of the semantic
(oracle/fcl_semantic_ref.py; the libraries are absent), so this is a measurement with a floor, not a parity claim."""
import numpy as np

from catgrasp_b200.synthetic import make_filter_case
from oracle import fcl_semantic_ref, filter_ref


def agreement(n_poses=256, res=0.0005):
    p1, p2, poses, sym, nocs_pose, c2n, g = make_filter_case(43, n_poses, 1)
    st, off, out = filter_ref.filter_ref(poses, sym, nocs_pose, c2n, g["gripper_in_grasp"], False, False, 0, g["open"], p1,
                                         g["enclosed"], p2)
    none = np.zeros((0, 3))
    _, _, unshifted = filter_ref.filter_ref(poses, sym, nocs_pose, c2n, g["gripper_in_grasp"], False, False, 0, g["open"], none,
                                            g["enclosed"], none)
    sdf_hit = st == 3
    sem_hit = np.zeros(len(poses), bool)
    for i in range(len(poses)):
        gic = unshifted[i].astype(np.float64) @ g["gripper_in_grasp"]
        sem_hit[i] = (fcl_semantic_ref.mesh_hits_points(g["open"]["V"], g["open"]["F"], gic, p1, res) or
                      fcl_semantic_ref.mesh_hits_points(g["enclosed"]["V"], g["enclosed"]["F"], gic, p2, res))
    return sdf_hit, sem_hit


def test_sdf_predicate_agrees_with_the_mesh_voxel_semantic():
    sdf_hit, sem_hit = agreement()
    agree = (sdf_hit == sem_hit).mean()
    print(f"agreement {agree:.3f}; SDF-only hits {(sdf_hit & ~sem_hit).sum()}, mesh/voxel-only hits {(~sdf_hit & sem_hit).sum()}, "
          f"both {(sdf_hit & sem_hit).sum()}, neither {(~sdf_hit & ~sem_hit).sum()} of {len(sdf_hit)}")
    assert sdf_hit.any() and (~sdf_hit).any()
    assert agree >= 0.9
