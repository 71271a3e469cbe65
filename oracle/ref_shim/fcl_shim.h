// Stand-in for the FCL types my_cpp/collision_manager.h names (FCL is not in /root/reference and not installed).
// Only what the reference HEADER needs to compile; the CollisionManager methods themselves are provided by
// collision_manager_sdf.cpp with the gripper-SDF predicate.  ORACLE / test infrastructure.
#pragma once
#include <Eigen/Dense>
#include <memory>
#include <vector>

namespace fcl {
template <typename S> using Vector3 = Eigen::Matrix<S, 3, 1>;
template <typename S> using Matrix3 = Eigen::Matrix<S, 3, 3>;
using Vector3f = Vector3<float>;     // common.cpp relies on `using namespace fcl` for the bare name Matrix3f
using Matrix3f = Matrix3<float>;
struct OBBRSSf {};
template <typename BV> class BVHModel {};

struct ShimGeometry {
  int kind = 0;                 // 0: gripper mesh (stands for its SDF), 1: point set
  int sdf_slot = -1;            // which registered SDF a mesh maps to
  std::vector<float> pts;       // kind 1: (P,3) row-major
};

template <typename S>
class CollisionObject {
 public:
  std::shared_ptr<ShimGeometry> geom;
  float pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // row-major 4x4
  void setTransform(const Matrix3<S> &R, const Vector3<S> &t) {
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) pose[r * 4 + c] = R(r, c);
      pose[r * 4 + 3] = t(r);
    }
  }
};
}  // namespace fcl
