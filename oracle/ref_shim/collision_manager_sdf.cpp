// CollisionManager methods for the oracle/_ref build of the reference's my_cpp/common.cpp.
//
// The reference implements them with FCL + octomap (my_cpp/collision_manager.cpp:15-111), neither of which is in
// /root/reference or installed, so my_cpp cannot be built as shipped.  This file keeps the reference's class
// (collision_manager.h is included from where it lies) and its argument checks, and answers isAnyCollision() with the
// gripper-SDF predicate of oracle/filter_ref.c (gripper_hits_ref): "some registered point lies inside the posed
// gripper".  Everything else in the .so -- filterGraspPose's pose arithmetic (Eigen), loop order, offset search,
// OpenMP structure -- is the reference's own translation unit, unmodified.  ORACLE / test infrastructure only.
#include "collision_manager.h"

#include <mutex>

extern "C" int gripper_hits_ref(const float *gripper_in_cam, const float *grid, const int *dims, const float *origin,
                                float res, int mode, const float *pts, int P);

namespace {
struct SdfSlot {
  int n_vertices = -1;
  const float *grid = nullptr;
  int dims[3] = {0, 0, 0};
  float origin[3] = {0, 0, 0};
  float res = 0.f;
};
SdfSlot g_slots[2];
int g_mode = 0;
}  // namespace

// The harness says which SDF stands for which mesh (keyed by the mesh's vertex count) before calling filterGraspPose.
extern "C" void ref_register_gripper_sdf(int slot, int n_vertices, const float *grid, const int *dims, const float *origin,
                                         float res) {
  SdfSlot &s = g_slots[slot];
  s.n_vertices = n_vertices;
  s.grid = grid;
  for (int i = 0; i < 3; i++) { s.dims[i] = dims[i]; s.origin[i] = origin[i]; }
  s.res = res;
}
extern "C" void ref_set_sdf_mode(int mode) { g_mode = mode; }

CollisionManager::CollisionManager() {}
CollisionManager::~CollisionManager() {}

int CollisionManager::registerMesh(Eigen::Ref<const Eigen::MatrixXf> V, Eigen::Ref<const Eigen::MatrixXi> F) {
  if (V.cols() != 3) { printf("vertices shape wrong: %dx%d\n", (int)V.rows(), (int)V.cols()); exit(1); }
  if (F.cols() != 3) { printf("faces shape wrong: %dx%d\n", (int)F.rows(), (int)F.cols()); exit(1); }
  CollisionObject<float> obj;
  obj.geom = std::make_shared<ShimGeometry>();
  obj.geom->kind = 0;
  for (int s = 0; s < 2; s++)
    if (g_slots[s].n_vertices == (int)V.rows()) obj.geom->sdf_slot = s;
  if (obj.geom->sdf_slot < 0) { printf("no SDF registered for a mesh with %d vertices\n", (int)V.rows()); exit(1); }
  int ob_id = _obs.size();
  _obs.push_back(obj);
  return ob_id;
}

int CollisionManager::registerPointCloud(Eigen::Ref<const Eigen::MatrixXf> pts, const float resolution) {
  if (pts.cols() != 3) { printf("point cloud shape wrong: %dx%d\n", (int)pts.rows(), (int)pts.cols()); exit(1); }
  (void)resolution;   // the SDF predicate tests the points themselves, not res-sized voxels around them
  CollisionObject<float> obj;
  obj.geom = std::make_shared<ShimGeometry>();
  obj.geom->kind = 1;
  obj.geom->pts.resize((size_t)pts.rows() * 3);
  for (int i = 0; i < pts.rows(); i++)
    for (int k = 0; k < 3; k++) obj.geom->pts[(size_t)i * 3 + k] = pts(i, k);
  int ob_id = _obs.size();
  _obs.push_back(obj);
  return ob_id;
}

void CollisionManager::setTransform(Eigen::Ref<const Eigen::MatrixXf> pose, const int ob_id) {
  if (pose.rows() != 4 || pose.cols() != 4) { printf("pose shape wrong: %dx%d\n", (int)pose.rows(), (int)pose.cols()); exit(1); }
  _obs[ob_id].setTransform(pose.block(0, 0, 3, 3), pose.block(0, 3, 3, 1));
}

bool CollisionManager::isAnyCollision() {
  for (size_t i = 0; i < _obs.size(); i++)
    for (size_t j = i + 1; j < _obs.size(); j++) {
      const auto &oi = _obs[i];
      const auto &oj = _obs[j];
      const auto *mesh = oi.geom->kind == 0 ? &oi : (oj.geom->kind == 0 ? &oj : nullptr);
      const auto *cloud = oi.geom->kind == 1 ? &oi : (oj.geom->kind == 1 ? &oj : nullptr);
      if (!mesh || !cloud) continue;
      const SdfSlot &s = g_slots[mesh->geom->sdf_slot];
      if (gripper_hits_ref(mesh->pose, s.grid, s.dims, s.origin, s.res, g_mode, cloud->geom->pts.data(),
                           (int)(cloud->geom->pts.size() / 3)))
        return true;
    }
  return false;
}
