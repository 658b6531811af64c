// Lock-free three-slot hand-off between the camera thread and the detector thread.
// Interface of the reference's TripleBuffer (reference include/irmv_detection/triple_buffer.hpp:15-49):
// the producer never waits, the consumer waits only for a fresh frame and always gets the newest
// one (older ones are dropped).  The slots here are the engine's pinned host frame buffers.
#pragma once
#include <array>
#include <atomic>

namespace irmv_detection
{
template <typename Buffer>
class TripleBuffer
{
public:
  explicit TripleBuffer(std::array<Buffer, 3> & slots) : back_(&slots[0]), middle_(&slots[1]), front_(&slots[2]) {}

  // slot the producer may fill right now
  Buffer * get_producer_buffer() { return back_; }

  // publish the filled slot; what was in the middle (possibly an unconsumed frame) becomes the
  // next slot to fill
  void producer_commit()
  {
    back_ = middle_.exchange(back_, std::memory_order_acq_rel);
    fresh_.store(true, std::memory_order_release);
    fresh_.notify_one();
  }

  // block until something new was published, then take it
  Buffer * get_consumer_buffer()
  {
    fresh_.wait(false, std::memory_order_acquire);
    front_ = middle_.exchange(front_, std::memory_order_acq_rel);
    fresh_.store(false, std::memory_order_release);
    return front_;
  }

private:
  Buffer * back_;                    // producer-owned
  std::atomic<Buffer *> middle_;     // exchanged by both sides
  Buffer * front_;                   // consumer-owned
  std::atomic<bool> fresh_{false};
};
}  // namespace irmv_detection
