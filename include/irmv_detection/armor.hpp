// Data types at both ends of the per-frame pipeline: detector class ids, light bars, armors.
// Same names and members as the reference (reference include/irmv_detection/armor.hpp:7-77);
// the Light constructor is written for the compat RotatedRect too.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>

#include "irmv_detection/cv_compat.hpp"

namespace irmv_detection
{
// 14 detector classes + UNKNOWN (class ids 0..13 of the network, 14 for anything else)
enum class ArmorClass { B1, B2, B3, B4, B5, BO, BS, R1, R2, R3, R4, R5, RO, RS, UNKNOWN };

enum class ArmorSize { SMALL, LARGE, UNKNOWN };

inline const char * armor_class_name(ArmorClass c)
{
  static constexpr const char * kNames[] = {"B1", "B2", "B3", "B4", "B5", "BO", "BS", "R1",
                                            "R2", "R3", "R4", "R5", "RO", "RS", "UNKNOWN"};
  const int i = static_cast<int>(c);
  return kNames[(i < 0 || i > 14) ? 14 : i];
}

inline ArmorClass armor_class_from_id(int id)
{
  return (id >= 0 && id < 14) ? static_cast<ArmorClass>(id) : ArmorClass::UNKNOWN;
}

struct Light : public cv::RotatedRect
{
  Light() = default;
#if IRMV_HAVE_OPENCV
  explicit Light(cv::RotatedRect box) : cv::RotatedRect(box)
  {
    std::array<cv::Point2f, 4> p;
    box.points(p.data());
    std::sort(p.begin(), p.end(), [](const cv::Point2f & a, const cv::Point2f & b) { return a.y < b.y; });
    set_ends((p[0] + p[1]) / 2, (p[2] + p[3]) / 2, cv::norm(p[0] - p[1]));
  }
#endif
  // top/bottom end points of the bar and its width, the three quantities PnP and the light
  // filter use
  Light(cv::Point2f top_pt, cv::Point2f bottom_pt, double bar_width) { set_ends(top_pt, bottom_pt, bar_width); }

  bool is_light(float min_ratio, float max_ratio, float max_angle) const
  {
    const double ratio = width / length;
    return min_ratio < ratio && ratio < max_ratio && tilt_angle < max_angle;
  }

  void offset_bbox(float min_x, float min_y)
  {
    center.x += min_x; center.y += min_y;
    top.x += min_x; top.y += min_y;
    bottom.x += min_x; bottom.y += min_y;
  }

  cv::Point2f top;
  cv::Point2f bottom;
  double length = 0;
  double width = 0;
  double tilt_angle = 0;

private:
  void set_ends(cv::Point2f t, cv::Point2f b, double w)
  {
    top = t; bottom = b; width = w;
    center = (t + b) / 2.0f;
    length = cv::norm(t - b);
    tilt_angle = std::atan2(std::abs(t.x - b.x), std::abs(t.y - b.y)) / 3.14159265358979323846 * 180.0;
  }
};

struct Armor
{
  Armor() = default;
  Armor(const Light & l1, const Light & l2)
  {
    const bool first_is_left = l1.center.x < l2.center.x;
    left_light = first_is_left ? l1 : l2;
    right_light = first_is_left ? l2 : l1;
    center = (left_light.center + right_light.center) / 2.0f;
  }

  Light left_light;
  Light right_light;
  ArmorSize size = ArmorSize::UNKNOWN;
  ArmorClass armor_class = ArmorClass::UNKNOWN;
  float confidence = 0.f;
  cv::Point2f center;
};
}  // namespace irmv_detection
