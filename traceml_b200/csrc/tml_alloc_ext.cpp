// tml_alloc_ext.cpp -- reads c10's CUDACachingAllocator counters directly.
//
// The reference pays three Python allocator-stat calls per step
// (torch.cuda.reset_peak_memory_stats / max_memory_allocated / max_memory_reserved,
// src/traceml/utils/step_memory.py:57,73-74), each of which builds a ~100-entry
// dict under the allocator mutex.  This extension returns the two counters as
// plain integers; the commit kernel then reads them from the host-mapped
// counter page (csrc/tml_engine.cu: k_commit).
#include <c10/cuda/CUDACachingAllocator.h>
#include <torch/extension.h>

#include <tuple>

namespace {

std::tuple<int64_t, int64_t> peak_bytes(int64_t device) {
  const auto st = c10::cuda::CUDACachingAllocator::getDeviceStats(static_cast<c10::DeviceIndex>(device));
  return {st.allocated_bytes[0].peak, st.reserved_bytes[0].peak};
}

void reset_peaks(int64_t device) {
  c10::cuda::CUDACachingAllocator::resetPeakStats(static_cast<c10::DeviceIndex>(device));
}

std::tuple<int64_t, int64_t> current_bytes(int64_t device) {
  const auto st = c10::cuda::CUDACachingAllocator::getDeviceStats(static_cast<c10::DeviceIndex>(device));
  return {st.allocated_bytes[0].current, st.reserved_bytes[0].current};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("peak_bytes", &peak_bytes, "(peak allocated, peak reserved) bytes of one device");
  m.def("reset_peaks", &reset_peaks, "reset the allocator's peak counters");
  m.def("current_bytes", &current_bytes, "(allocated, reserved) bytes right now");
}
