// YoloEngine: same public interface as the reference class
// (reference include/irmv_detection/yolo_engine.hpp:16-73), implemented over libirmv_b200.so.
// TensorRT / NPP / unified memory are gone: the object owns an irmv_engine handle (C ABI,
// include/irmv_cabi.h) whose pinned-host frame slot is what get_src_image_buffer() returns.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "irmv_cabi.h"
#include "irmv_detection/armor.hpp"

namespace irmv_detection
{
class YoloEngine
{
public:
  struct bbox
  {
    std::array<float, 4> xyxy;
    float score;
    ArmorClass class_id;

    bool operator==(const bbox & other) const
    {
      return xyxy == other.xyxy && score == other.score && class_id == other.class_id;
    }
  };

  // onnx_file_path: the reference swaps the extension for ".engine"; here it is swapped for ".irmw"
  YoloEngine(const std::string & onnx_file_path, cv::Size src_image_size, bool enable_profiling = false);
  ~YoloEngine();
  YoloEngine(const YoloEngine &) = delete;
  YoloEngine & operator=(const YoloEngine &) = delete;

  std::vector<bbox> detect();
  void visualize_bboxes(cv::Mat & image, const std::vector<bbox> & bboxes) const;
  double get_profiling_time() const { return inference_time_ms_; }
  const cv::Mat & get_rotated_image() const;
  uint8_t * get_src_image_buffer() const { return src_image_buffer_; }
  // not in the reference class: the C-ABI handle, for the stages that fuse into this engine's replay
  // (ArmorExtractor::enable, irmv_engine_enable_pnp)
  irmv_engine * handle() const { return engine_; }

private:
  irmv_engine * engine_ = nullptr;
  cv::Size src_image_size_;
  uint8_t * src_image_buffer_ = nullptr;     // pinned host slot, stable for the object's life
  mutable cv::Mat rotated_image_;            // header over the engine's pinned rotated-frame buffer
  bool enable_profiling_ = false;
  double inference_time_ms_ = 0.0;
};
}  // namespace irmv_detection
