// tml_step_ext.cpp -- torch-side glue of the STEP PATH (python module traceml_b200._tml_step).
//
// Two jobs, both plumbing between PyTorch and the C-ABI of libtraceml_b200.so:
//
//  1. Allocator counters.  The reference pays three Python allocator-stat calls per
//     step (torch.cuda.reset_peak_memory_stats / max_memory_allocated /
//     max_memory_reserved, src/traceml/utils/step_memory.py:57,73-74), each building
//     a ~100-entry dict under the allocator mutex.  Here c10's DeviceStats is read
//     directly and the two integers go straight into tml_step_commit, which hands
//     them to the commit kernel through the host-mapped counter page.
//
//  2. Region open / close without Python-side stream lookups or ctypes marshalling:
//     begin()/end()/host()/commit() resolve torch's CURRENT stream natively
//     (c10::cuda::getCurrentCUDAStream) and call the C-ABI entry points, which are
//     bound once with dlsym -- the boundary stays the C-ABI, this file adds no
//     telemetry logic of its own.
//  3. The native process sampler (BASELINE config 4: 1 kHz telemetry).  The reference's
//     ProcessSampler runs in a Python thread and fights the training thread for the GIL
//     (src/traceml/samplers/process_sampler.py:208-238, runtime/runtime.py:110-140); here
//     a C++ thread reads process CPU time (CLOCK_PROCESS_CPUTIME_ID), RSS
//     (/proc/self/statm) and the allocator's current counters, and commits one 64-B
//     ProcRecord per period through tml_proc_commit on its own non-blocking stream.
#include <c10/cuda/CUDACachingAllocator.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <torch/extension.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>

#include "../../include/traceml_b200.h"

namespace {

using begin_fn = int (*)(void*, uint32_t, void*);
using end_fn = int (*)(void*, uint32_t, int, void*);
using host_fn = int (*)(void*, uint32_t, uint64_t);
using commit_fn = int (*)(void*, uint64_t, uint64_t, uint64_t, uint32_t, double, void*);
using proc_commit_fn = int (*)(void*, const tml_proc_record*, void*);

void* g_ctx = nullptr;
int g_device = 0;
begin_fn g_begin = nullptr;
end_fn g_end = nullptr;
host_fn g_host = nullptr;
commit_fn g_commit = nullptr;
proc_commit_fn g_proc_commit = nullptr;

std::thread g_sampler;
std::atomic<bool> g_sampler_run{false};
std::atomic<uint64_t> g_sampler_seq{0};
std::atomic<uint64_t> g_sampler_late{0};

// Last line of defence (the package registers an atexit that unbinds first): a joinable
// std::thread destroyed at static destruction is std::terminate, which would mask the run's
// real exit code.  Declared after g_sampler, so it is destroyed before it.
struct SamplerJoin {
  ~SamplerJoin() {
    if (g_sampler_run.exchange(false) && g_sampler.joinable()) g_sampler.join();
    else if (g_sampler.joinable()) g_sampler.join();
  }
} g_sampler_join;

void bind(const std::string& lib_path, uint64_t ctx, int64_t device) {
  void* h = dlopen(lib_path.c_str(), RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
  g_begin = reinterpret_cast<begin_fn>(dlsym(h, "tml_phase_begin"));
  g_end = reinterpret_cast<end_fn>(dlsym(h, "tml_phase_end"));
  g_host = reinterpret_cast<host_fn>(dlsym(h, "tml_phase_host"));
  g_commit = reinterpret_cast<commit_fn>(dlsym(h, "tml_step_commit"));
  g_proc_commit = reinterpret_cast<proc_commit_fn>(dlsym(h, "tml_proc_commit"));
  if (!g_begin || !g_end || !g_host || !g_commit || !g_proc_commit)
    throw std::runtime_error("libtraceml_b200.so lacks a step-path symbol");
  g_ctx = reinterpret_cast<void*>(ctx);
  g_device = static_cast<int>(device);
}

void unbind() {
  if (g_sampler_run.exchange(false) && g_sampler.joinable()) g_sampler.join();
  g_ctx = nullptr;
}

inline void* cur_stream() {
  return c10::cuda::getCurrentCUDAStream(static_cast<c10::DeviceIndex>(g_device)).stream();
}

int64_t begin(int64_t phase) {
  if (!g_ctx) return -3;
  return g_begin(g_ctx, static_cast<uint32_t>(phase), cur_stream());
}

int64_t end(int64_t phase, int64_t slot) {
  if (!g_ctx) return -3;
  return g_end(g_ctx, static_cast<uint32_t>(phase), static_cast<int>(slot), cur_stream());
}

int64_t host(int64_t phase, int64_t dur_ns) {
  if (!g_ctx) return -3;
  return g_host(g_ctx, static_cast<uint32_t>(phase), dur_ns < 0 ? 0ull : static_cast<uint64_t>(dur_ns));
}

// mem_device >= 0: read that device's allocator peaks and flag the record HAS_MEM
int64_t commit(int64_t step, int64_t mem_device, double host_ts) {
  if (!g_ctx) return -3;
  uint64_t a = 0, r = 0;
  uint32_t flags = 0;
  if (mem_device >= 0) {
    const auto st = c10::cuda::CUDACachingAllocator::getDeviceStats(static_cast<c10::DeviceIndex>(mem_device));
    a = static_cast<uint64_t>(st.allocated_bytes[0].peak);
    r = static_cast<uint64_t>(st.reserved_bytes[0].peak);
    flags = 1u;
  }
  return g_commit(g_ctx, static_cast<uint64_t>(step), a, r, flags, host_ts, cur_stream());
}

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

// ---------------------------------------------------------------- native process sampler
double now_s(clockid_t id) {
  timespec ts;
  clock_gettime(id, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void sampler_loop(int64_t period_us, int64_t start_seq) {
  void* ctx = g_ctx;
  const int dev = g_device;
  cudaSetDevice(dev);
  cudaStream_t stream = nullptr;
  cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const long page = sysconf(_SC_PAGESIZE);
  const long cores = sysconf(_SC_NPROCESSORS_ONLN);
  const int fd = open("/proc/self/statm", O_RDONLY);
  double last_wall = now_s(CLOCK_MONOTONIC), last_cpu = now_s(CLOCK_PROCESS_CPUTIME_ID);
  uint64_t seq = (uint64_t)start_seq;
  auto next = std::chrono::steady_clock::now();
  const auto period = std::chrono::microseconds(period_us);
  while (g_sampler_run.load(std::memory_order_relaxed)) {
    next += period;
    std::this_thread::sleep_until(next);
    if (std::chrono::steady_clock::now() > next + period) {  // fell behind: resynchronise
      g_sampler_late.fetch_add(1);
      next = std::chrono::steady_clock::now();
    }
    const double wall = now_s(CLOCK_MONOTONIC), cpu = now_s(CLOCK_PROCESS_CPUTIME_ID);
    tml_proc_record r;
    r.seq = ++seq;
    r.ts = now_s(CLOCK_REALTIME);
    r.cpu_pct = wall > last_wall ? 100.0 * (cpu - last_cpu) / (wall - last_wall) : 0.0;
    last_wall = wall; last_cpu = cpu;
    r.rss = 0;
    if (fd >= 0) {
      char buf[128];
      const ssize_t k = pread(fd, buf, sizeof(buf) - 1, 0);
      if (k > 0) {
        buf[k] = 0;
        unsigned long size = 0, resident = 0;
        if (sscanf(buf, "%lu %lu", &size, &resident) == 2) r.rss = (uint64_t)resident * (uint64_t)page;
      }
    }
    const auto st = c10::cuda::CUDACachingAllocator::getDeviceStats(static_cast<c10::DeviceIndex>(dev));
    r.mem_alloc = (uint64_t)st.allocated_bytes[0].current;
    r.mem_resv = (uint64_t)st.reserved_bytes[0].current;
    r.mem_total = (uint64_t)total_b;
    r.flags = TML_PROC_GPU_AVAILABLE | TML_PROC_HAS_GPU_METRICS;
    r.cpu_cores = (uint32_t)(cores > 0 ? cores : 0);
    g_proc_commit(ctx, &r, stream);
    g_sampler_seq.store(seq);
  }
  cudaStreamSynchronize(stream);
  cudaStreamDestroy(stream);
  if (fd >= 0) close(fd);
}

void sampler_start(int64_t period_us, int64_t start_seq) {
  if (!g_ctx) throw std::runtime_error("bind() first");
  if (g_sampler_run.exchange(true)) return;
  g_sampler = std::thread(sampler_loop, period_us < 50 ? 50 : period_us, start_seq);
}

std::tuple<int64_t, int64_t> sampler_stop() {
  if (g_sampler_run.exchange(false) && g_sampler.joinable()) g_sampler.join();
  return {(int64_t)g_sampler_seq.load(), (int64_t)g_sampler_late.load()};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("bind", &bind, "bind the C-ABI step-path entry points and the engine context");
  m.def("unbind", &unbind);
  m.def("begin", &begin, "tml_phase_begin on torch's current stream -> slot");
  m.def("end", &end, "tml_phase_end on torch's current stream");
  m.def("host", &host, "tml_phase_host");
  m.def("commit", &commit, "tml_step_commit with the allocator peaks of mem_device (or none if < 0)");
  m.def("peak_bytes", &peak_bytes, "(peak allocated, peak reserved) bytes of one device");
  m.def("reset_peaks", &reset_peaks, "reset the allocator's peak counters");
  m.def("current_bytes", &current_bytes, "(allocated, reserved) bytes right now");
  m.def("sampler_start", &sampler_start, "start the native process sampler (period in us, first seq)");
  m.def("sampler_stop", &sampler_stop, py::call_guard<py::gil_scoped_release>(),
        "stop it -> (last seq, periods missed)");
}
