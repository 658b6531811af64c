// Camera side of the hand-off: the types the detector node exchanges with its cameras and a
// VirtualCamera with the reference's constructor and threading
// (reference include/irmv_detection/camera.hpp:14-62, src/camera.cpp:9-93).
//
// The reference's VirtualCamera decodes a video file with cv::VideoCapture; OpenCV is not part of
// this build, so the source here is a raw frame file (frames of image_size.width x height x 3 bytes
// back to back, looped) or frames handed over in memory.  Everything downstream is the reference's
// protocol: the stream thread fills the producer slot of a TripleBuffer over the three
// caller-owned buffers (the engines' pinned host slots, reference src/irm_detector.cpp:68-72),
// stamps and commits it at `fps`; the receive thread waits for a fresh slot and runs the callback
// on it (drop-old, never block the producer, reference README.md:60-63).  MVCamera (vendor SDK,
// reference src/mv_camera.cpp) is out of scope.
#pragma once
#include <array>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "irmv_detection/cv_compat.hpp"
#include "irmv_detection/triple_buffer.hpp"

namespace irmv_detection
{
class Camera
{
public:
  struct Config
  {
    double exposure_time = 0;
    int analog_gain = 0;
    int saturation = 0;
    int gamma = 0;
    cv::Size image_size = cv::Size(0, 0);
    std::array<uint8_t *, 3> image_buffers{};
  };
  struct StampedImage
  {
    cv::Mat image;
    std::chrono::time_point<std::chrono::system_clock> time_stamp;
    int id = 0;
  };
  class invalid_camera_error : public std::runtime_error
  {
  public:
    using runtime_error::runtime_error;
  };
  using CameraCallback = std::function<void(StampedImage &)>;
  Camera() = default;
  virtual ~Camera() = default;
};

class VirtualCamera : public Camera
{
public:
  // video_path: raw frame file (see above).  Throws invalid_camera_error when it cannot be read and
  // std::invalid_argument when its size is not a whole number of frames, like the reference does for
  // an unreadable video / a size mismatch (src/camera.cpp:15-22).
  explicit VirtualCamera(const Config & config, const std::string & video_path, const CameraCallback & callback, int fps = 100)
  : config_(config), camera_callback_(callback), fps_(fps)
  {
    FILE * f = fopen(video_path.c_str(), "rb");
    if (!f) throw invalid_camera_error("Cannot open video file");
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    const size_t fb = frame_bytes();
    if (fb == 0 || bytes <= 0 || static_cast<size_t>(bytes) % fb != 0) {
      fclose(f);
      throw std::invalid_argument("Image size does not match");
    }
    frames_.resize(static_cast<size_t>(bytes));
    const size_t got = fread(frames_.data(), 1, frames_.size(), f);
    fclose(f);
    if (got != frames_.size()) throw invalid_camera_error("Cannot open video file");
    start();
  }

  // frames already in memory (n x frame_bytes)
  explicit VirtualCamera(const Config & config, std::vector<uint8_t> frames, const CameraCallback & callback, int fps = 100)
  : config_(config), camera_callback_(callback), fps_(fps), frames_(std::move(frames))
  {
    if (frame_bytes() == 0 || frames_.empty() || frames_.size() % frame_bytes() != 0)
      throw std::invalid_argument("Image size does not match");
    start();
  }

  ~VirtualCamera() override
  {
    shutdown_ = true;
    if (stream_thread_.joinable()) stream_thread_.join();
    triple_buffer_->producer_commit();      // commit a dummy to wake up the consumer thread
    if (receive_thread_.joinable()) receive_thread_.join();
  }

  long frames_produced() const { return produced_.load(); }
  long frames_consumed() const { return consumed_.load(); }

private:
  size_t frame_bytes() const { return static_cast<size_t>(config_.image_size.width) * config_.image_size.height * 3; }

  void start()
  {
    for (int i = 0; i < 3; i++) {
      if (!config_.image_buffers[i]) throw std::invalid_argument("Image buffers are not set");
      stamped_img_buf_[i].image = cv::Mat(config_.image_size, CV_8UC3, config_.image_buffers[i]);
      stamped_img_buf_[i].id = i;
    }
    triple_buffer_ = std::make_unique<TripleBuffer<StampedImage>>(stamped_img_buf_);
    stream_thread_ = std::thread(&VirtualCamera::stream_thread, this);
    receive_thread_ = std::thread(&VirtualCamera::receive_thread, this);
  }

  void stream_thread()
  {
    namespace chrono = std::chrono;
    const auto interval = chrono::duration_cast<chrono::system_clock::duration>(chrono::duration<double>(1.0 / fps_));
    const size_t fb = frame_bytes(), nframes = frames_.size() / fb;
    size_t next = 0;
    while (!shutdown_) {
      const auto start_time = chrono::system_clock::now();
      StampedImage * slot = triple_buffer_->get_producer_buffer();
      memcpy(slot->image.data, frames_.data() + next * fb, fb);      // the "retrieve" of the reference (src/camera.cpp:49)
      next = (next + 1) % nframes;
      slot->time_stamp = start_time;
      triple_buffer_->producer_commit();
      produced_++;
      std::this_thread::sleep_until(start_time + interval);
    }
  }

  void receive_thread()
  {
    while (!shutdown_) {
      StampedImage * slot = triple_buffer_->get_consumer_buffer();
      if (shutdown_) break;
      camera_callback_(*slot);
      consumed_++;
    }
  }

  Config config_;
  CameraCallback camera_callback_;
  int fps_;
  std::vector<uint8_t> frames_;
  std::array<StampedImage, 3> stamped_img_buf_;
  std::unique_ptr<TripleBuffer<StampedImage>> triple_buffer_;
  std::atomic<bool> shutdown_{false};
  std::atomic<long> produced_{0}, consumed_{0};
  std::thread stream_thread_;
  std::thread receive_thread_;
};
}  // namespace irmv_detection
