// C++ host side of the drop-in: irmv_detection::YoloEngine over the C ABI.
// Behaviour follows reference src/yolo_engine.cpp: constructor builds everything and warms up
// (:24-117), detect() = launch + sync + parse with optional wall-clock profiling (:153-177),
// a missing model file prints a message and exits (:37-40).
#include "irmv_detection/yolo_engine.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <iostream>

namespace irmv_detection
{
namespace fs = std::filesystem;

YoloEngine::YoloEngine(const std::string & onnx_file_path, cv::Size src_image_size, bool enable_profiling)
: src_image_size_(src_image_size), enable_profiling_(enable_profiling)
{
  fs::path weights(onnx_file_path);
  weights.replace_extension(".irmw");
  if (!fs::exists(weights)) {
    std::cout << "Please build the weight file (" << weights.string() << ") first." << std::endl;
    exit(0);
  }
  irmv_engine_config cfg;
  irmv_engine_config_default(&cfg);
  cfg.src_width = src_image_size.width;
  cfg.src_height = src_image_size.height;
  cfg.max_batch = 1;
  cfg.num_slots = 1;                        // the node makes one engine per ring slot
  if (irmv_engine_create(weights.string().c_str(), &cfg, &engine_) != 0) {
    std::cerr << "[YoloEngine] " << irmv_last_error() << std::endl;
    exit(1);
  }
  src_image_buffer_ = irmv_engine_src_buffer(engine_, 0);
  for (int i = 0; i < 50; i++) detect();    // warm up, like the reference
}

YoloEngine::~YoloEngine() { irmv_engine_destroy(engine_); }

std::vector<YoloEngine::bbox> YoloEngine::detect()
{
  std::chrono::high_resolution_clock::time_point t0;
  if (enable_profiling_) t0 = std::chrono::high_resolution_clock::now();
  irmv_bbox raw[1024];
  int n = 0;
  std::vector<bbox> out;
  if (irmv_engine_detect(engine_, 0, raw, 1024, &n) != 0) {
    std::cerr << "[YoloEngine::detect] " << irmv_last_error() << std::endl;
    return out;
  }
  out.reserve(n);
  for (int i = 0; i < n; i++) {
    bbox b;
    b.xyxy = {raw[i].xyxy[0], raw[i].xyxy[1], raw[i].xyxy[2], raw[i].xyxy[3]};
    b.score = raw[i].score;
    b.class_id = armor_class_from_id(raw[i].class_id);
    out.emplace_back(b);
  }
  if (enable_profiling_) {
    inference_time_ms_ =
      std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
  }
  return out;
}

const cv::Mat & YoloEngine::get_rotated_image() const
{
  // a view of the engine's pinned, address-stable rotated-frame buffer (the reference's cv::Mat is a
  // view of the buffer its graph rotates in place, yolo_engine.hpp:34): after the first call every
  // detect() refreshes the buffer, so the per-frame call of the node (src/irm_detector.cpp:183)
  // copies nothing and allocates nothing
  const uint8_t * view = nullptr;
  if (irmv_engine_rotated_view(engine_, 0, &view) != 0) {
    std::cerr << "[YoloEngine::get_rotated_image] " << irmv_last_error() << std::endl;
    return rotated_image_;
  }
  if (rotated_image_.data != view) {
    rotated_image_ = cv::Mat(
      cv::Size(src_image_size_.width, src_image_size_.height), CV_8UC3, const_cast<uint8_t *>(view));
  }
  return rotated_image_;
}

void YoloEngine::visualize_bboxes(cv::Mat & image, const std::vector<bbox> & bboxes) const
{
  if (image.cols != src_image_size_.width || image.rows != src_image_size_.height) {
    std::cerr << "[YoloEngine::visualize_bboxes] Image size mismatch" << std::endl;
    return;
  }
#if IRMV_HAVE_OPENCV
  for (const auto & b : bboxes) {
    const bool blue = armor_class_name(b.class_id)[0] == 'B';
    const cv::Scalar color = blue ? cv::Scalar(0, 0, 255) : cv::Scalar(255, 0, 0);
    cv::rectangle(image, cv::Point(b.xyxy[0], b.xyxy[1]), cv::Point(b.xyxy[2], b.xyxy[3]), color, 2);
    cv::putText(image, armor_class_name(b.class_id), cv::Point(b.xyxy[0], b.xyxy[1]), cv::FONT_HERSHEY_SIMPLEX, 1, color, 2);
  }
#else
  (void)bboxes;   // drawing needs OpenCV imgproc; debug-only path of the node
#endif
}
}  // namespace irmv_detection
