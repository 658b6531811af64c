// VirtualCamera + TripleBuffer hand-off without a GPU: a producer at a fixed rate, a consumer that is
// slower than the producer for a while (frames must be dropped, never reordered, never torn), the three
// caller-owned buffers, ids and time stamps as the detector node sees them
// (reference src/camera.cpp:9-93, src/irm_detector.cpp:68-72,176-183).
#include <chrono>
#include <cstdio>
#include <mutex>
#include <thread>
#include <vector>

#include "irmv_detection/camera.hpp"

using namespace irmv_detection;

int main()
{
  const int W = 64, H = 48, N = 50;
  const size_t fb = static_cast<size_t>(W) * H * 3;
  // frame k is filled with the byte k + 1, so a torn frame shows as two values in one buffer
  std::vector<uint8_t> frames(fb * N);
  for (int k = 0; k < N; k++) memset(frames.data() + k * fb, k + 1, fb);
  std::array<std::vector<uint8_t>, 3> bufs;
  Camera::Config cfg;
  cfg.image_size = cv::Size(W, H);
  for (int i = 0; i < 3; i++) { bufs[i].assign(fb, 0); cfg.image_buffers[i] = bufs[i].data(); }

  std::mutex mu;
  std::vector<int> seen;
  std::vector<int> ids;
  bool torn = false, bad_buffer = false, bad_time = false;
  auto last_stamp = std::chrono::system_clock::time_point::min();
  int calls = 0;
  auto callback = [&](Camera::StampedImage & img) {
    std::lock_guard<std::mutex> lock(mu);
    const uint8_t v = img.image.data[0];
    for (size_t i = 0; i < fb; i += 997) if (img.image.data[i] != v) torn = true;
    if (img.image.data != cfg.image_buffers[img.id]) bad_buffer = true;
    if (img.time_stamp < last_stamp) bad_time = true;
    last_stamp = img.time_stamp;
    seen.push_back(v);
    ids.push_back(img.id);
    if (++calls <= 20) std::this_thread::sleep_for(std::chrono::milliseconds(12));   // slower than the 500 fps producer
  };
  long produced = 0, consumed = 0;
  {
    VirtualCamera cam(cfg, frames, callback, 500);
    std::this_thread::sleep_for(std::chrono::milliseconds(600));
    produced = cam.frames_produced();
    consumed = cam.frames_consumed();
  }
  printf("produced %ld consumed %ld callbacks %zu\n", produced, consumed, seen.size());
  if (torn) { printf("FAIL: torn frame\n"); return 1; }
  if (bad_buffer) { printf("FAIL: image does not live in the caller's buffer of its id\n"); return 1; }
  if (bad_time) { printf("FAIL: time stamps went backwards\n"); return 1; }
  if (seen.size() < 30 || produced <= static_cast<long>(seen.size())) { printf("FAIL: expected dropped frames\n"); return 1; }
  // frames come in production order (modulo the loop of N frames): the step between two consumed frames is
  // 1..N-1 forward, never 0 (the same frame twice) while the producer is running
  int drops = 0;
  for (size_t i = 1; i < seen.size(); i++) {
    const int step = ((seen[i] - seen[i - 1]) % N + N) % N;
    if (step == 0) { printf("FAIL: frame %d delivered twice\n", seen[i]); return 1; }
    if (step > 1) drops += step - 1;
  }
  printf("dropped %d frames while the consumer was slow\n", drops);
  if (drops == 0) { printf("FAIL: no drops seen\n"); return 1; }
  // a missing file and a size mismatch fail like the reference
  bool threw = false;
  try { VirtualCamera c2(cfg, std::string("/nonexistent/video.raw"), callback, 100); } catch (const Camera::invalid_camera_error &) { threw = true; }
  if (!threw) { printf("FAIL: missing file accepted\n"); return 1; }
  threw = false;
  try { VirtualCamera c3(cfg, std::vector<uint8_t>(fb + 1), callback, 100); } catch (const std::invalid_argument &) { threw = true; }
  if (!threw) { printf("FAIL: size mismatch accepted\n"); return 1; }
  printf("VIRTUAL_CAMERA_OK\n");
  return 0;
}
