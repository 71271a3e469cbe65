"""Build recipe for oracle/_ref/libmycpp_ref.so: the REFERENCE's own my_cpp/common.cpp (filterGraspPose,
directionVecToRotation, augmentGraspPoses, get_ik_within_limits, makeOccupancyGridFromCloudScan) and its generated
ikfast solver, compiled from the sources where they lie under /root/reference with the reference's own options
(my_cpp/CMakeLists.txt:5-6: Release, -std=c++14 -fopenmp, no -march) against the Eigen 3.2.92 vendored in the reference.

What is NOT the reference in that library (my_cpp as shipped is unbuildable here: FCL, octomap, Boost and pybind11's
Python headers are neither vendored nor installed):
  * oracle/ref_shim/fcl/**, fcl_shim.h     type stand-ins so collision_manager.h compiles;
  * oracle/ref_shim/collision_manager_sdf.cpp  the CollisionManager methods, answering isAnyCollision() with the
    gripper-SDF predicate of filter_ref.c instead of FCL's mesh-vs-octree test;
  * oracle/ref_shim/octomap/octomap.h      a restatement of the few octomap calls common.cpp:324-431 makes;
  * oracle/ref_shim/{pybind11,boost}/**    empty headers (included by common.h, nothing used).
So the library pins the POSE LOGIC and control flow (Eigen arithmetic, approach test, offset search, survivor output) of
filterGraspPose and the flow of makeOccupancyGridFromCloudScan; the FCL / octomap boundary itself stays unpinned.

Runs only where /root/reference exists (the authoring container).  Output goes to oracle/_ref/ (git-ignored, travels
to the GPU box with the snapshot).  No reference source is copied into the repository.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libmycpp_ref.so")
IKFAST = f"{REF}/ikfast_pybind/src/kuka_iiwa14/ikfast0x1000004a.Transform6D.0_1_3_4_5_6_f2.cpp"


def available():
    return os.path.exists(f"{REF}/my_cpp/common.cpp")


def build(force=False):
    if not available():
        return LIB if os.path.exists(LIB) else None
    shim = os.path.join(HERE, "ref_shim")
    mine = [os.path.join(shim, "collision_manager_sdf.cpp"), os.path.join(shim, "ref_api.cpp"),
            os.path.join(HERE, "filter_ref.c"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in mine):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    eig = f"{REF}/PointGroup/lib/pointgroup_ops"
    inc = ["-I", shim, "-I", f"{eig}/eigen3", "-I", eig, "-I", f"{REF}/my_cpp", "-I", f"{REF}/ikfast_pybind/src"]
    cxx = ["g++", "-O3", "-DNDEBUG", "-std=c++14", "-fopenmp", "-fPIC", "-w"] + inc
    objs = []
    for name, src, extra in (("common", f"{REF}/my_cpp/common.cpp", []),
                             ("ikfast", IKFAST, ["-DIKFAST_NO_MAIN", "-DIKFAST_HAS_LIBRARY"]),
                             ("cm_sdf", mine[0], []), ("ref_api", mine[1], [])):
        o = os.path.join(OUT_DIR, name + ".o")
        subprocess.check_call(cxx + extra + ["-c", src, "-o", o])
        objs.append(o)
    o = os.path.join(OUT_DIR, "filter_ref.o")      # same options as oracle/build_oracle.py: identical predicate arithmetic
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-fPIC", "-c", mine[2], "-o", o])
    objs.append(o)
    subprocess.check_call(["g++", "-shared", "-fopenmp", "-o", LIB] + objs + ["-lm"])
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
