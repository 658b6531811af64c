// PnPSolver: same public interface as the reference class
// (reference include/irmv_detection/pnp_solver.hpp:12-38); cv::solvePnP(IPPE) on the CPU is
// replaced by the batched CUDA IPPE kernel behind irmv_pnp_solve (include/irmv_cabi.h).
#pragma once
#include <array>
#include <vector>

#include "irmv_cabi.h"
#include "irmv_detection/armor.hpp"

namespace irmv_detection
{
class PnPSolver
{
public:
  PnPSolver(const std::array<double, 9> & camera_matrix, const std::vector<double> & distortion_coefficients);
  ~PnPSolver();
  PnPSolver(const PnPSolver &) = delete;
  PnPSolver & operator=(const PnPSolver &) = delete;

  // rvec/tvec come back as 3x1 CV_64F, like cv::solvePnP leaves them (read with .at<double>)
  bool solvePnP(const Armor & armor, cv::Mat & rvec, cv::Mat & tvec) const;

  // Distance between the armor centre and the principal point (the reference reads its CV_64F
  // camera matrix as float there, src/pnp_solver.cpp:56-57; this is the intended computation)
  float calculateDistanceToCenter(const cv::Point2f & image_point);

private:
  irmv_pnp * solver_ = nullptr;
};
}  // namespace irmv_detection
