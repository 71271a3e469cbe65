"""CPU ORACLE (test infrastructure only): the SEMANTIC of the reference's geometry predicate
(my_cpp/collision_manager.cpp:15-111) -- "a posed triangle mesh intersects an occupied octree voxel".

FCL and octomap are not in /root/reference and not installed (versions unpinned), so this is a restatement from their
documented behaviour, NOT a pin: occupied voxels = unique keys floor(x / res) of the registered points (one
``updateNode(point, true)`` makes a leaf occupied), a voxel = the axis-aligned cube of side ``res`` centred at
(key + 0.5) * res, collision = some cube overlaps some posed triangle (separating-axis test with the 13 axes of a
box/triangle pair, float64).  It exists to MEASURE how often the gripper-SDF predicate the product uses
gives the same verdict as the mesh-vs-voxel predicate the reference uses -- see tests/test_fcl_semantic_agreement.py.
"""
import numpy as np


def occupied_voxel_centers(points, res):
    keys = np.unique(np.floor(np.asarray(points, np.float64) / res).astype(np.int64), axis=0)
    return (keys + 0.5) * res


def _axis_separates(v, axis, half):
    """v (V,3,3) triangle vertices relative to the box centres, axis (V,3) or (3,), half: box half extent."""
    p = np.einsum("vij,vj->vi", v, np.broadcast_to(axis, (v.shape[0], 3)))
    r = half * np.abs(np.broadcast_to(axis, (v.shape[0], 3))).sum(1)
    return (p.min(1) > r) | (p.max(1) < -r)


def any_box_triangle_overlap(centers, half, tris):
    """centers (V,3), half: scalar half side, tris (T,3,3) -> True iff some box overlaps some triangle."""
    eye = np.eye(3)
    for tri in tris:
        v = tri[None, :, :] - centers[:, None, :]                     # (V,3,3)
        alive = np.ones(len(centers), bool)
        for a in range(3):                                            # box face normals
            alive &= ~((v[:, :, a].min(1) > half) | (v[:, :, a].max(1) < -half))
        if not alive.any():
            continue
        v = v[alive]
        e = np.stack([tri[1] - tri[0], tri[2] - tri[1], tri[0] - tri[2]])
        n = np.cross(e[0], e[1])
        sep = _axis_separates(v, n, half)                             # triangle plane
        for a in range(3):
            for k in range(3):
                if sep.all():
                    break
                sep |= _axis_separates(v, np.cross(eye[a], e[k]), half)
        if (~sep).any():
            return True
    return False


def mesh_hits_points(V, F, gripper_in_cam, points, res):
    """The reference's isAnyCollision() semantic for one posed mesh and one registered point set."""
    T = np.asarray(gripper_in_cam, np.float64)
    Vc = (T[:3, :3] @ np.asarray(V, np.float64).T).T + T[:3, 3]
    tris = Vc[np.asarray(F, np.int64)]                                # (T,3,3)
    centers = occupied_voxel_centers(points, res)
    # cheap cull: voxels farther than the mesh's bounding sphere cannot touch it
    c = Vc.mean(0)
    rad = np.linalg.norm(Vc - c, axis=1).max() + res
    centers = centers[np.linalg.norm(centers - c, axis=1) <= rad]
    if len(centers) == 0:
        return False
    return any_box_triangle_overlap(centers, res / 2.0, tris)
