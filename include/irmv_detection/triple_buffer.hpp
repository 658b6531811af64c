// Lock-free three-slot hand-off between the camera thread and the detector thread.
// Interface of the reference's TripleBuffer (reference include/irmv_detection/triple_buffer.hpp:15-49):
// the producer never waits, the consumer waits only for a fresh frame and always gets the newest
// one (older ones are dropped).  The slots here are the engine's pinned host frame buffers.
//
// Difference from the reference: it keeps the "new frame" flag in a second atomic that is set after
// the pointer exchange (triple_buffer.hpp:26-32) and cleared after the consumer's exchange
// (:34-40); a consumer that runs between the producer's exchange and its flag store first takes
// the new frame and then, woken by the late flag, takes the OLDER frame that it had just handed
// back (observed as "frame 808 after 811" under a full-speed producer).  Here slot index and
// freshness live in ONE atomic word, so a hand-off is a single exchange on either side and the
// consumer can only ever move to a newer frame.
#pragma once
#include <array>
#include <atomic>
#include <cstdint>

namespace irmv_detection
{
template <typename Buffer>
class TripleBuffer
{
public:
  explicit TripleBuffer(std::array<Buffer, 3> & slots) : slots_(slots.data()) {}

  // slot the producer may fill right now
  Buffer * get_producer_buffer() { return slots_ + back_; }

  // publish the filled slot; what was in the middle (possibly an unconsumed frame) becomes the
  // next slot to fill
  void producer_commit()
  {
    back_ = middle_.exchange(back_ | kFresh, std::memory_order_acq_rel) & kIndex;
    middle_.notify_one();
  }

  // block until something new was published, then take it
  Buffer * get_consumer_buffer()
  {
    for (uint32_t m = middle_.load(std::memory_order_acquire); !(m & kFresh); m = middle_.load(std::memory_order_acquire))
      middle_.wait(m, std::memory_order_acquire);
    // only the producer writes the middle word besides us, and it always sets kFresh
    front_ = middle_.exchange(front_, std::memory_order_acq_rel) & kIndex;
    return slots_ + front_;
  }

private:
  static constexpr uint32_t kFresh = 4, kIndex = 3;
  Buffer * slots_;
  uint32_t back_ = 0;                     // producer-owned slot index
  std::atomic<uint32_t> middle_{1};       // slot index | kFresh, exchanged by both sides
  uint32_t front_ = 2;                    // consumer-owned slot index
};
}  // namespace irmv_detection
