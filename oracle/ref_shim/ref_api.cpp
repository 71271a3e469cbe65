// extern "C" entry points into the oracle/_ref build of the reference's my_cpp/common.cpp (plain pointers, row-major
// float32 matrices), for oracle/mycpp_ref.py.  ORACLE / test infrastructure only.
#include <cstdio>
#include "common.h"
#include "collision_manager.h"

#include <cstring>

namespace {
Eigen::Matrix4f m4(const float *p) {
  Eigen::Matrix4f m;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m(r, c) = p[r * 4 + c];
  return m;
}
Eigen::MatrixXf mx(const float *p, int rows, int cols) {
  Eigen::MatrixXf m(rows, cols);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) m(r, c) = p[(size_t)r * cols + c];
  return m;
}
void put4(const Eigen::Matrix4f &m, float *p) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) p[r * 4 + c] = m(r, c);
}
}  // namespace

// verbose switch of the next ref_filterGraspPose calls: the reference then prints its four rejection counters
// (common.cpp:316-319) to stdout, which oracle/mycpp_ref.py captures
static int g_verbose = 0;
extern "C" void ref_set_verbose(int v) { g_verbose = v; }

// common.h:60.  Meshes are passed as vertex counts only (the shim keys the SDF on them).  Returns the survivor count;
// at most `cap` poses are written.
extern "C" int ref_filterGraspPose(const float *grasp_poses, int G, const float *symmetry_tfs, int S, const float *nocs_pose,
                                   const float *canonical_to_nocs, const float *cam_in_world, const float *ee_in_grasp,
                                   const float *gripper_in_grasp, int filter_dir, int filter_ik, int adjust,
                                   const double *upper, const double *lower, int n_joint, int open_nv, int encl_nv,
                                   const float *open_pts, int P1, const float *encl_pts, int P2, float octo_resolution,
                                   float *out_poses, int cap) {
  vectorMatrix4f gp, st;
  for (int i = 0; i < G; i++) gp.push_back(m4(grasp_poses + (size_t)i * 16));
  for (int j = 0; j < S; j++) st.push_back(m4(symmetry_tfs + (size_t)j * 16));
  std::vector<double> up(upper, upper + n_joint), lo(lower, lower + n_joint);
  Eigen::MatrixXf oV = Eigen::MatrixXf::Zero(open_nv, 3), eV = Eigen::MatrixXf::Zero(encl_nv, 3);
  Eigen::MatrixXi F = Eigen::MatrixXi::Zero(1, 3);
  vectorMatrix4f out = filterGraspPose(gp, st, m4(nocs_pose), m4(canonical_to_nocs), m4(cam_in_world), m4(ee_in_grasp),
                                       m4(gripper_in_grasp), filter_dir != 0, filter_ik != 0, adjust != 0, up, lo, oV, F, eV, F,
                                       mx(open_pts, P1, 3), mx(encl_pts, P2, 3), octo_resolution, g_verbose != 0);
  if (g_verbose) fflush(stdout);
  for (size_t i = 0; i < out.size() && (int)i < cap; i++) put4(out[i], out_poses + i * 16);
  return (int)out.size();
}

// common.h:56
extern "C" void ref_directionVecToRotation(const float *direction, const float *ref, float *out9) {
  Eigen::Matrix3f R = directionVecToRotation(Eigen::Vector3f(direction[0], direction[1], direction[2]),
                                             Eigen::Vector3f(ref[0], ref[1], ref[2]));
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) out9[r * 3 + c] = R(r, c);
}

// common.h:54
extern "C" int ref_ik_solution_count(const float *ee_in_base, const double *upper, const double *lower, int n_joint) {
  std::vector<double> up(upper, upper + n_joint), lo(lower, lower + n_joint);
  return (int)get_ik_within_limits(m4(ee_in_base), up, lo).size();
}

// common.h:61.  Returns the number of occupied samples; at most `cap` rows are written.
extern "C" int ref_makeOccupancyGridFromCloudScan(const float *pts, int P, const float *K, float resolution, float *out,
                                                  int cap) {
  Eigen::Matrix3f Km;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Km(r, c) = K[r * 3 + c];
  Eigen::MatrixXf o = makeOccupancyGridFromCloudScan(mx(pts, P, 3), Km, resolution);
  for (int i = 0; i < o.rows() && i < cap; i++)
    for (int k = 0; k < 3; k++) out[(size_t)i * 3 + k] = o(i, k);
  return (int)o.rows();
}
