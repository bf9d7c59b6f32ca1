// Library-wide state: per-thread error message, launch counter, version.
#include "common.cuh"

#include <cstdarg>

namespace b200 {

std::string& lastError() {
  static thread_local std::string msg;
  return msg;
}

int fail(int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  lastError() = buf;
  return status;
}

std::atomic<unsigned long long> g_launches{0};

} // namespace b200

extern "C" const char* bout_b200_last_error(void) { return b200::lastError().c_str(); }
extern "C" int bout_b200_version(void) { return 100; }
extern "C" unsigned long long bout_b200_launch_count(void) {
  return b200::g_launches.load(std::memory_order_relaxed);
}

// ---- device memory for host layers that do not link the CUDA runtime themselves -------------------
extern "C" int bout_b200_malloc(void** ptr, size_t bytes) {
  B200_REQUIRE(ptr != nullptr, BOUT_B200_ERR_ARG, "malloc: null argument");
  B200_CUDA_OK(cudaMalloc(ptr, bytes));
  B200_CUDA_OK(cudaMemset(*ptr, 0, bytes));
  return 0;
}
extern "C" int bout_b200_free(void* ptr) {
  B200_CUDA_OK(cudaFree(ptr));
  return 0;
}
extern "C" int bout_b200_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  B200_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, b200::asStream(stream)));
  B200_CUDA_OK(cudaStreamSynchronize(b200::asStream(stream)));
  return 0;
}
extern "C" int bout_b200_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  B200_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, b200::asStream(stream)));
  B200_CUDA_OK(cudaStreamSynchronize(b200::asStream(stream)));
  return 0;
}
extern "C" int bout_b200_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  B200_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, b200::asStream(stream)));
  return 0;
}
extern "C" int bout_b200_stream_sync(void* stream) {
  B200_CUDA_OK(cudaStreamSynchronize(b200::asStream(stream)));
  return 0;
}

// ---- streams and page-locked staging for the output step --------------------------------------
// The RK loop runs on the caller's stream; a dump snapshot is copied device -> host on a second, non-blocking
// stream into page-locked memory while the next steps compute (PhysicsModelMonitor::call,
// src/physics/physicsmodel.cxx:241-268, writes synchronously from host-resident fields).
extern "C" int bout_b200_stream_create(void** stream) {
  B200_REQUIRE(stream != nullptr, BOUT_B200_ERR_ARG, "stream_create: null argument");
  cudaStream_t s = nullptr;
  B200_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = reinterpret_cast<void*>(s);
  return 0;
}
extern "C" int bout_b200_stream_destroy(void* stream) {
  B200_REQUIRE(stream != nullptr, BOUT_B200_ERR_ARG, "stream_destroy: the default stream cannot be destroyed");
  B200_CUDA_OK(cudaStreamDestroy(b200::asStream(stream)));
  return 0;
}
// Work queued on `waiter` after this call starts only when everything queued on `signaller` so far has finished
extern "C" int bout_b200_stream_wait(void* waiter, void* signaller) {
  cudaEvent_t ev = nullptr;
  B200_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t err = cudaEventRecord(ev, b200::asStream(signaller));
  if (err == cudaSuccess) {
    err = cudaStreamWaitEvent(b200::asStream(waiter), ev, 0);
  }
  cudaEventDestroy(ev); // released by the runtime once the recorded work has completed
  if (err != cudaSuccess) {
    return b200::fail(BOUT_B200_ERR_CUDA, "stream_wait failed: %s", cudaGetErrorString(err));
  }
  return 0;
}
extern "C" int bout_b200_host_alloc(void** ptr, size_t bytes) {
  B200_REQUIRE(ptr != nullptr, BOUT_B200_ERR_ARG, "host_alloc: null argument");
  B200_CUDA_OK(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return 0;
}
extern "C" int bout_b200_host_free(void* ptr) {
  B200_CUDA_OK(cudaFreeHost(ptr));
  return 0;
}
// Queues the copy and returns; the caller synchronises `stream` before reading dst (dst from bout_b200_host_alloc)
extern "C" int bout_b200_memcpy_d2h_async(void* dst, const void* src, size_t bytes, void* stream) {
  B200_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, b200::asStream(stream)));
  return 0;
}
