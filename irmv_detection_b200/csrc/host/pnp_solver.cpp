// C++ host side of the drop-in: irmv_detection::PnPSolver over the C ABI
// (reference src/pnp_solver.cpp:7-59).
#include "irmv_detection/pnp_solver.hpp"

#include <iostream>

namespace irmv_detection
{
PnPSolver::PnPSolver(const std::array<double, 9> & camera_matrix, const std::vector<double> & dist_coeffs)
{
  double d[5] = {0, 0, 0, 0, 0};
  for (size_t i = 0; i < 5 && i < dist_coeffs.size(); i++) d[i] = dist_coeffs[i];
  if (irmv_pnp_create(camera_matrix.data(), d, 0, &solver_) != 0)
    std::cerr << "[PnPSolver] " << irmv_last_error() << std::endl;
}

PnPSolver::~PnPSolver() { irmv_pnp_destroy(solver_); }

bool PnPSolver::solvePnP(const Armor & armor, cv::Mat & rvec, cv::Mat & tvec) const
{
  // image points in the reference's order: left.bottom, left.top, right.top, right.bottom
  const float pts[8] = {armor.left_light.bottom.x, armor.left_light.bottom.y, armor.left_light.top.x,
                        armor.left_light.top.y,    armor.right_light.top.x,   armor.right_light.top.y,
                        armor.right_light.bottom.x, armor.right_light.bottom.y};
  double r[3], t[3];
  int ok = 0;
  if (!solver_ || irmv_pnp_solve(solver_, pts, r, t, &ok) != 0 || !ok) return false;
  rvec.create(3, 1, CV_64F);
  tvec.create(3, 1, CV_64F);
  for (int i = 0; i < 3; i++) {
    rvec.at<double>(i) = r[i];
    tvec.at<double>(i) = t[i];
  }
  return true;
}

float PnPSolver::calculateDistanceToCenter(const cv::Point2f & image_point)
{
  return irmv_pnp_distance_to_center(solver_, image_point.x, image_point.y);
}
}  // namespace irmv_detection
