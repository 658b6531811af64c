// IrmDetector: header-compatible with the reference node class
// (reference include/irmv_detection/irm_detector.hpp:19-84) where ROS2 is installed, plus the node's
// per-frame data path as a ROS-free class (IrmDetectorCore) that the node body -- and the tests here --
// drive.  The ROS2 shell itself (parameters, publishers, RViz markers, component registration,
// reference src/irm_detector.cpp:25-174,357-411) is out of scope: this image has
// no rclcpp, so only the declaration is kept compatible; INTEGRATION.md shows the three edits that turn
// the reference's irm_detector.cpp into a caller of IrmDetectorCore.
#pragma once
#include <array>
#include <memory>
#include <vector>

#include "irmv_cabi.h"
#include "irmv_detection/armor.hpp"
#include "irmv_detection/armor_extractor.hpp"
#include "irmv_detection/camera.hpp"
#include "irmv_detection/pnp_solver.hpp"
#include "irmv_detection/yolo_engine.hpp"

namespace irmv_detection
{
// What message_callback publishes per armor (auto_aim_interfaces/Armor fields filled at
// reference src/irm_detector.cpp:213-230), without the ROS message type.
struct ArmorPose
{
  double position[3];      // armor_msg.pose.position      = tvec
  double orientation[4];   // armor_msg.pose.orientation   = tf2 quaternion (x, y, z, w) of Rodrigues(rvec)
  float distance_to_image_center;
  ArmorClass armor_class;
  ArmorSize size;
};

// IrmDetector::message_callback (reference src/irm_detector.cpp:176-245) without ROS: three engines (one per
// triple-buffer slot, :35-38,68-72), detect -> extract_armors -> solvePnP -> Rodrigues -> quaternion ->
// distance, all inside the engine's CUDA graph replay; the host only reads the per-armor payload.
class IrmDetectorCore
{
public:
  // camera_matrix / dist_coeffs as CameraInfoManager delivers them (reference :42-52); corner_scale maps
  // source pixels to the calibration frame (1, 1 when the calibration matches the camera resolution).
  IrmDetectorCore(
    const std::string & onnx_file_path, cv::Size image_input_size, const std::array<double, 9> & camera_matrix,
    const std::vector<double> & dist_coeffs, float corner_scale_x = 1.f, float corner_scale_y = 1.f)
  {
    for (auto & e : yolo_engines_) {
      e = std::make_unique<YoloEngine>(onnx_file_path, image_input_size, false);
      extractor_.enable(*e);
      double D[5] = {0, 0, 0, 0, 0};
      for (size_t i = 0; i < 5 && i < dist_coeffs.size(); i++) D[i] = dist_coeffs[i];
      irmv_engine_enable_pnp(e->handle(), camera_matrix.data(), D, corner_scale_x, corner_scale_y);
    }
  }

  // the three source buffers the camera writes (reference src/irm_detector.cpp:68-72)
  std::array<uint8_t *, 3> image_buffers() const
  {
    return {yolo_engines_[0]->get_src_image_buffer(), yolo_engines_[1]->get_src_image_buffer(),
            yolo_engines_[2]->get_src_image_buffer()};
  }

  ArmorExtractor & extractor() { return extractor_; }

  // One frame: image.id names the slot / engine that holds it.  Returns the armors with a pose.
  std::vector<ArmorPose> message_callback(Camera::StampedImage & image)
  {
    YoloEngine & engine = *yolo_engines_[image.id];
    const std::vector<YoloEngine::bbox> bboxes = engine.detect();
    std::vector<ArmorPose> out;
    if (bboxes.empty()) return out;
    poses_.resize(1024);
    armors_.resize(1024);
    if (irmv_engine_fetch_armor_poses(engine.handle(), -1, 1, poses_.data()) != 0) return out;
    if (irmv_engine_fetch_armors(engine.handle(), -1, 1, armors_.data()) != 0) return out;
    for (size_t i = 0; i < bboxes.size(); i++) {
      if (!poses_[i].ok) continue;              // no armor in this box, or solvePnP failed (reference :207-209)
      ArmorPose p;
      for (int k = 0; k < 3; k++) p.position[k] = poses_[i].position[k];
      for (int k = 0; k < 4; k++) p.orientation[k] = poses_[i].orientation[k];
      p.distance_to_image_center = poses_[i].distance_to_image_center;
      p.armor_class = armor_class_from_id(armors_[i].class_id);
      p.size = armors_[i].size ? ArmorSize::LARGE : ArmorSize::SMALL;
      out.push_back(p);
    }
    return out;
  }

private:
  std::array<std::unique_ptr<YoloEngine>, 3> yolo_engines_;
  ArmorExtractor extractor_;
  std::vector<irmv_pose> poses_;
  std::vector<irmv_armor> armors_;
};
}  // namespace irmv_detection

#if __has_include(<rclcpp/rclcpp.hpp>)
#include <rclcpp/rclcpp.hpp>
namespace irmv_detection
{
// Same public surface as the reference class (include/irmv_detection/irm_detector.hpp:23-27); the body
// (src/irm_detector.cpp) keeps its ROS2 plumbing and forwards message_callback to IrmDetectorCore.
class IrmDetector
{
public:
  explicit IrmDetector(const rclcpp::NodeOptions & options);
  rclcpp::node_interfaces::NodeBaseInterface::SharedPtr get_node_base_interface() const
  {
    return node_->get_node_base_interface();
  }

private:
  void message_callback(Camera::StampedImage & image);
  rclcpp::Node::SharedPtr node_;
  std::unique_ptr<IrmDetectorCore> core_;
  std::unique_ptr<Camera> camera_;
};
}  // namespace irmv_detection
#endif
