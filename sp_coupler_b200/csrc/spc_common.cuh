// Shared host/device helpers of the spcpl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "spcpl_b200.h"

struct spc_ctx {
  int device;
  int num_sms;
  int max_smem_optin;
  uint32_t magic;
  // tuning overrides; always 0 (= production choice) unless set through libspcpl_b200_tune.so (-DSPC_TUNING)
  int k1_variant, ijk_variant, k2_threads, k3_threads, proj_threads;
};
#define SPC_MAGIC 0x53504342u

namespace spc {

// physical constants, splib/sputils.py:14-19
constexpr double pref0 = 1.0e5;
constexpr double rd = 287.04;
constexpr double rv = 461.5;
constexpr double cp = 1004.0;
constexpr double rlv = 2.53e6;
constexpr double grav = 9.81;

void set_error(const char* fmt, ...);
int k1_configure(spc_ctx* c);   // slab_reduce.cu: opt the streaming kernels in to their dynamic shared memory on c->device
int check_handle(spc_handle h);
int cuda_fail(cudaError_t e, const char* what);

#define SPC_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      spc::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

#define SPC_CUDA(call)                                      \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return spc::cuda_fail(e__, #call); \
  } while (0)

// Enters the handle's device for the duration of a call, restores the caller's device after.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
      switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// ---- device helpers -------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double ldd(const T* p, size_t i) {
  return (double)__ldg(p + i);
}
template <typename T>
__device__ __forceinline__ void std_(T* p, size_t i, double v) {
  if (p) p[i] = (T)v;
}

// numpy.searchsorted(a, v, side="right") on an ascending array: first index with a[i] > v
__device__ __forceinline__ int upper_bound(const double* a, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// numpy.searchsorted(a, v, side="left"): first index with a[i] >= v
__device__ __forceinline__ int lower_bound(const double* a, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// numpy.interp for one abscissa given its bracket j = upper_bound(xp, x) - 1 (sputils.py:82-86).
// Same expression as numpy's C loop: slope*(x - xp[j]) + fp[j]; compiled with --fmad=false so the
// float64 result is bit-identical to numpy's.
__device__ __forceinline__ double interp_at(const double* xp, const double* fp, int n, double x, int j) {
  if (j < 0) return fp[0];
  if (j >= n - 1) return fp[n - 1];
  if (xp[j] == x) return fp[j];
  double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  return slope * (x - xp[j]) + fp[j];
}

}  // namespace spc
