// ArmorExtractor: IrmDetector::extract_armors (reference src/irm_detector.cpp:292-355, a private
// member of the ROS2 node) as a stand-alone class over libirmv_b200.so.  Members carry the node's
// parameter names and defaults (reference src/irm_detector.cpp:152,162-173,
// include/irmv_detection/irm_detector.hpp:50-58).
//
//   extract_armors(image, bboxes)   the reference's signature: `image` is get_rotated_image(), a
//                                   packed u8x3 host image; one C-ABI call (irmv_extract_armors)
//   extract_armors(engine, bboxes)  the B200 form: the engine already ran the stage inside its replay
//                                   (enable(engine) once), this only reads the armors back
#pragma once
#include <vector>

#include "irmv_cabi.h"
#include "irmv_detection/armor.hpp"
#include "irmv_detection/yolo_engine.hpp"

namespace irmv_detection
{
class ArmorExtractor
{
public:
  int binary_threshold_ = 150;
  double light_min_ratio_ = 0.1;
  double light_max_ratio_ = 0.4;
  double light_max_angle_ = 40.0;
  double armor_min_small_center_distance_ = 0.8;
  double armor_max_small_center_distance_ = 3.2;
  double armor_min_large_center_distance_ = 3.2;
  double armor_max_large_center_distance_ = 5.5;

  std::vector<Armor> extract_armors(const cv::Mat & image, const std::vector<YoloEngine::bbox> & bboxes) const;

  // Fuse the stage into `engine`'s replay (between NMS and PnP) with the current parameters.
  bool enable(YoloEngine & engine) const;
  // Armors of the engine's last detect(); `bboxes` is what that detect() returned.
  std::vector<Armor> extract_armors(const YoloEngine & engine, const std::vector<YoloEngine::bbox> & bboxes) const;

private:
  irmv_armor_params params() const;
};
}  // namespace irmv_detection
