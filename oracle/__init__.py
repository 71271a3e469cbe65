"""CPU oracle for the CaTGrasp hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``catgrasp_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, as the checker and as the timed CPU baseline -- never as the product.

Pinning status (DESIGN.md section "Oracle"):
  * pointnet_ref.py, pn2_ref.py : PINNED against the reference's own pointnet2.py executed in the
    authoring container (tests/golden/make_golden.py -> tests/golden/*.npz).
  * transforms_ref.py, aligning_ref.py : PINNED against the reference's own predicter.py / dataset_grasp.py /
    dataset_nunocs.py / augmentations.py / aligning.py / Utils.load_model executed in the authoring container with
    the absent third-party imports stubbed (tests/golden/make_golden_hostpath.py -> tests/golden/host_*.npz).
  * filter_ref.c                : pose logic restates my_cpp/common.cpp; the FCL/octomap geometry
    predicate cannot be built here -> parity unpinned (SDF predicate per meshpy/sdf.py instead).
"""
