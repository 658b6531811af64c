// C++ host side of the drop-in: IrmDetector::extract_armors (reference src/irm_detector.cpp:292-355)
// over the C ABI.  The per-box OpenCV chain of the reference (cvtColor, threshold, findContours,
// minAreaRect) runs as one CUDA kernel launch for all boxes (csrc/armors.cu).
#include "irmv_detection/armor_extractor.hpp"

#include <iostream>

namespace irmv_detection
{
namespace
{
std::vector<Armor> to_armors(const std::vector<irmv_armor> & raw, size_t n)
{
  std::vector<Armor> armors;
  for (size_t i = 0; i < n; i++) {
    const irmv_armor & a = raw[i];
    if (!a.valid) continue;
    // pts = left.bottom, left.top, right.top, right.bottom (reference src/pnp_solver.cpp:41-44)
    Light left(cv::Point2f(a.pts[2], a.pts[3]), cv::Point2f(a.pts[0], a.pts[1]), 0.0);
    Light right(cv::Point2f(a.pts[4], a.pts[5]), cv::Point2f(a.pts[6], a.pts[7]), 0.0);
    Armor armor;
    armor.left_light = left;
    armor.right_light = right;
    armor.center = cv::Point2f(a.center[0], a.center[1]);
    armor.size = a.size ? ArmorSize::LARGE : ArmorSize::SMALL;
    armor.armor_class = armor_class_from_id(a.class_id);
    armor.confidence = a.score;
    armors.emplace_back(armor);
  }
  return armors;
}
}  // namespace

irmv_armor_params ArmorExtractor::params() const
{
  irmv_armor_params p;
  irmv_armor_params_default(&p);
  p.binary_threshold = binary_threshold_;
  p.light_min_ratio = static_cast<float>(light_min_ratio_);     // Light::is_light takes floats (armor.hpp:31)
  p.light_max_ratio = static_cast<float>(light_max_ratio_);
  p.light_max_angle = static_cast<float>(light_max_angle_);
  p.min_small_center_distance = armor_min_small_center_distance_;
  p.max_small_center_distance = armor_max_small_center_distance_;
  p.min_large_center_distance = armor_min_large_center_distance_;
  p.max_large_center_distance = armor_max_large_center_distance_;
  return p;
}

std::vector<Armor> ArmorExtractor::extract_armors(const cv::Mat & image, const std::vector<YoloEngine::bbox> & bboxes) const
{
  if (bboxes.empty() || image.empty()) return {};
  std::vector<irmv_bbox> boxes(bboxes.size());
  for (size_t i = 0; i < bboxes.size(); i++) {
    for (int k = 0; k < 4; k++) boxes[i].xyxy[k] = bboxes[i].xyxy[k];
    boxes[i].score = bboxes[i].score;
    boxes[i].class_id = static_cast<int>(bboxes[i].class_id);
  }
  const int count = static_cast<int>(boxes.size());
  std::vector<irmv_armor> raw(boxes.size());
  const irmv_armor_params p = params();
  // `image` is already the rotated view, in its own channel order: no rotation, bytes as they are
  if (irmv_extract_armors(image.data, 0, 1, image.cols, image.rows, IRMV_CH_PASSTHROUGH, 0, boxes.data(), &count, count,
                          &p, 0, raw.data()) != 0) {
    std::cerr << "[ArmorExtractor::extract_armors] " << irmv_last_error() << std::endl;
    return {};
  }
  return to_armors(raw, raw.size());
}

bool ArmorExtractor::enable(YoloEngine & engine) const
{
  const irmv_armor_params p = params();
  if (irmv_engine_enable_armors(engine.handle(), &p) != 0) {
    std::cerr << "[ArmorExtractor::enable] " << irmv_last_error() << std::endl;
    return false;
  }
  return true;
}

std::vector<Armor> ArmorExtractor::extract_armors(const YoloEngine & engine, const std::vector<YoloEngine::bbox> & bboxes) const
{
  std::vector<irmv_armor> raw(1024);
  if (irmv_engine_fetch_armors(engine.handle(), -1, 1, raw.data()) != 0) {
    std::cerr << "[ArmorExtractor::extract_armors] " << irmv_last_error() << std::endl;
    return {};
  }
  return to_armors(raw, bboxes.size() < raw.size() ? bboxes.size() : raw.size());
}
}  // namespace irmv_detection
