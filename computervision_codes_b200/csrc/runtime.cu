// Library runtime: error strings, device queries.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace tcn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};  // kernels of this library enqueued (or captured into a graph) so far

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return TCN_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return TCN_OK;
}

long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) v = (getenv("TCN_NO_PDL") == nullptr) ? 1 : 0;
  return v == 1;
}

// Programmatic dependent launch is used inside captured graphs only.  Round 2 measured rare stale reads (1-40 frames of
// one layer's input) with the early trigger the kernels carried then (common.cuh: TCN_PDL_TRIGGER); they were far more
// frequent on eager launches (3-30 % of first forwards submitted while earlier work was still draining) than under
// graph replay.  The trigger is gone; eager launches keep a full stream dependency on top of that -- they are bound by
// the host's dispatch anyway.  TCN_PDL_EAGER=1 lifts the restriction for experiments (tools/exp/test_hunt.py).
bool pdl_allowed_on(cudaStream_t stream) {
  if (!pdl_enabled()) return false;
  static int eager = -1;
  if (eager < 0) eager = (getenv("TCN_PDL_EAGER") != nullptr) ? 1 : 0;
  if (eager == 1) return true;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return st == cudaStreamCaptureStatusActive;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    cached = prop.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace tcn

extern "C" int tcn_version(void) { return TCN_VERSION; }
extern "C" const char* tcn_last_error(void) { return tcn::g_err; }
extern "C" long long tcn_launch_count(void) { return tcn::launch_count(); }

extern "C" int tcn_device_info(int* cc_major, int* cc_minor, int* nsm, int* built_for_sm) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    tcn::set_error("tcn_device_info: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
    cudaGetLastError();
    return TCN_ERR_CUDA;
  }
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (nsm) *nsm = prop.multiProcessorCount;
  if (built_for_sm) *built_for_sm = 100;
  return TCN_OK;
}
