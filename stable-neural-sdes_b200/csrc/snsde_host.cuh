// Host-side plumbing shared by every C-ABI entry point (include/snsde.h): the thread-local error string,
// a device guard that restores the caller's current device, and the exception fence of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <exception>
#include <new>

#include "../../include/snsde.h"

namespace snsde {

// Formats the thread-local message returned by snsde_last_error() and returns `code` (defined in snsde_api.cu).
int fail(int code, const char* fmt, ...);

// cudaSetDevice(dev) for the lifetime of the object; the caller's device is restored on every exit path, so an
// engine call on a non-current device never changes PyTorch's current device behind its back.
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err != cudaSuccess) { prev = -1; return; }
    if (prev != dev) err = cudaSetDevice(dev);
    else prev = -1;                       // nothing to restore
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace snsde

// No C++ exception crosses the C ABI: every extern "C" body sits between these two macros.
#define SNSDE_API_BEGIN try {
#define SNSDE_API_END(on_error)                                                                       \
  } catch (const std::bad_alloc&) {                                                                   \
    snsde::fail(SNSDE_ERR_INTERNAL, "host allocation failed (std::bad_alloc)");                      \
    return on_error;                                                                                  \
  } catch (const std::exception& e) {                                                                 \
    snsde::fail(SNSDE_ERR_INTERNAL, "unexpected C++ exception: %s", e.what());                       \
    return on_error;                                                                                  \
  } catch (...) {                                                                                     \
    snsde::fail(SNSDE_ERR_INTERNAL, "unexpected C++ exception");                                     \
    return on_error;                                                                                  \
  }
