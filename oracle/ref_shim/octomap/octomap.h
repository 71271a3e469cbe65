// Stand-in for the octomap API my_cpp/common.cpp:324-431 uses (octomap is not in /root/reference and not installed).
// A restatement of octomap's published behaviour (OcTreeBaseImpl::coordToKey / keyToCoord, OccupancyOcTreeBase::
// insertPointCloud with the default hit/miss log-odds, OccupancyOcTreeBase::castRay, octomath::Vector3) on a hash set of
// occupied leaf keys.  ORACLE / test infrastructure; versions of the real library are not pinned by the reference.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <unordered_set>
#include <vector>

namespace octomath {
class Vector3 {
 public:
  Vector3(float x = 0.f, float y = 0.f, float z = 0.f) { data[0] = x; data[1] = y; data[2] = z; }
  const float &operator()(unsigned i) const { return data[i]; }
  float &operator()(unsigned i) { return data[i]; }
  float x() const { return data[0]; }
  float y() const { return data[1]; }
  float z() const { return data[2]; }
  double norm_sq() const { return (x() * x() + y() * y() + z() * z()); }
  double norm() const { return std::sqrt(norm_sq()); }
  Vector3 &operator/=(float v) {
    v = 1.0f / v;
    for (unsigned i = 0; i < 3; i++) data[i] *= v;
    return *this;
  }
  Vector3 &normalize() {
    double len = norm();
    if (len > 0) *this /= (float)len;
    return *this;
  }
  Vector3 normalized() const {
    Vector3 r(*this);
    r.normalize();
    return r;
  }

 private:
  float data[3];
};
}  // namespace octomath

namespace octomap {
typedef octomath::Vector3 point3d;

class Pointcloud {
 public:
  void push_back(const point3d &p) { pts.push_back(p); }
  size_t size() const { return pts.size(); }
  std::vector<point3d> pts;
};

class OcTree {
 public:
  explicit OcTree(double res) : resolution(res), resolution_factor(1.0 / res) {}

  bool coordToKeyChecked(double c, unsigned &key) const {
    int scaled = ((int)std::floor(resolution_factor * c)) + (int)tree_max_val;
    if (scaled < 0 || (unsigned)scaled >= 2 * tree_max_val) return false;
    key = (unsigned)scaled;
    return true;
  }
  double keyToCoord(unsigned key) const { return (double((int)key - (int)tree_max_val) + 0.5) * resolution; }

  // One call with default parameters: every end point's leaf becomes occupied (log-odds +0.85 > 0); cells only traversed
  // by rays become free.  castRay(..., ignoreUnknown=true) treats free and unknown alike, so only the occupied set is kept.
  void insertPointCloud(const Pointcloud &scan, const point3d & /*sensor_origin*/, double /*maxrange*/ = -1.,
                        bool /*lazy_eval*/ = false, bool /*discretize*/ = false) {
    for (const auto &p : scan.pts) {
      unsigned k[3];
      if (coordToKeyChecked(p(0), k[0]) && coordToKeyChecked(p(1), k[1]) && coordToKeyChecked(p(2), k[2])) occ.insert(pack(k));
    }
  }

  bool castRay(const point3d &origin, const point3d &directionP, point3d &end, bool ignoreUnknownCells = false,
               double maxRange = -1.0) const {
    unsigned cur[3];
    for (int i = 0; i < 3; i++)
      if (!coordToKeyChecked(origin(i), cur[i])) return false;
    if (occ.count(pack(cur))) {
      end = point3d((float)keyToCoord(cur[0]), (float)keyToCoord(cur[1]), (float)keyToCoord(cur[2]));
      return true;
    } else if (!ignoreUnknownCells) {
      end = point3d((float)keyToCoord(cur[0]), (float)keyToCoord(cur[1]), (float)keyToCoord(cur[2]));
      return false;
    }
    point3d direction = directionP.normalized();
    const bool max_range_set = (maxRange > 0.0);
    int step[3];
    double tMax[3], tDelta[3];
    for (unsigned i = 0; i < 3; ++i) {
      if (direction(i) > 0.0) step[i] = 1;
      else if (direction(i) < 0.0) step[i] = -1;
      else step[i] = 0;
      if (step[i] != 0) {
        double voxelBorder = keyToCoord(cur[i]);
        voxelBorder += double(step[i] * resolution * 0.5);
        tMax[i] = (voxelBorder - origin(i)) / direction(i);
        tDelta[i] = resolution / std::fabs(direction(i));
      } else {
        tMax[i] = std::numeric_limits<double>::max();
        tDelta[i] = std::numeric_limits<double>::max();
      }
    }
    if (step[0] == 0 && step[1] == 0 && step[2] == 0) return false;
    const double maxrange_sq = maxRange * maxRange;
    for (;;) {
      unsigned dim;
      if (tMax[0] < tMax[1]) dim = (tMax[0] < tMax[2]) ? 0 : 2;
      else dim = (tMax[1] < tMax[2]) ? 1 : 2;
      if ((step[dim] < 0 && cur[dim] == 0) || (step[dim] > 0 && cur[dim] == 2 * tree_max_val - 1)) return false;
      cur[dim] += step[dim];
      tMax[dim] += tDelta[dim];
      end = point3d((float)keyToCoord(cur[0]), (float)keyToCoord(cur[1]), (float)keyToCoord(cur[2]));
      if (max_range_set) {
        double d2 = 0.0;
        for (unsigned j = 0; j < 3; j++) d2 += ((end(j) - origin(j)) * (end(j) - origin(j)));
        if (d2 > maxrange_sq) return false;
      }
      if (occ.count(pack(cur))) return true;
      // free or unknown: with ignoreUnknownCells the walk continues; the reference always passes true
      if (!ignoreUnknownCells) return false;
    }
  }

  double resolution, resolution_factor;
  static constexpr unsigned tree_max_val = 32768;

 private:
  static uint64_t pack(const unsigned k[3]) { return ((uint64_t)k[0] << 32) | ((uint64_t)k[1] << 16) | (uint64_t)k[2]; }
  std::unordered_set<uint64_t> occ;
};
}  // namespace octomap
