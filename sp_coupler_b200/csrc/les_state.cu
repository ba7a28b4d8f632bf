// spc_set_les_state: profile -> volume broadcast with uniform noise (spcpl.set_les_state,
// splib/spcpl.py:274-294), HBM-write-bound. Noise comes from counter-based Philox4x32-10 so that
// the volume is a pure function of (seed, stream, global column, element) — identical on any
// sharding and reproducible on the host (sp_coupler_b200/synth.py: les_state_volume).
// Compiled with --fmad=false: prof + amp*n - sub is evaluated exactly as numpy does.
#include <math.h>

#include "spc_common.cuh"

namespace {

constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kM0, c.x), lo0 = kM0 * c.x;
    const uint32_t hi1 = __umulhi(kM1, c.z), lo1 = kM1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += kW0;
    k1 += kW1;
  }
  return c;
}

__device__ __forceinline__ double noise(uint32_t x) {
  const double u = (double)(x >> 8) * (1.0 / 16777216.0);
  return 2.0 * u - 1.0;
}

struct K5Args {
  const double *prof, *sub;
  void* vol;
  double amp;
  uint32_t stream_id, seed;
  int col0, clamp0, ncol, nk;
  long long S, nE, ngrp;  // slab elements, elements per column, Philox groups per column
};

template <typename T>
__device__ __forceinline__ double value(const K5Args& a, int c, long long e, uint32_t x) {
  const int k = (int)(e / a.S);
  double v = __ldg(a.prof + (size_t)c * a.nk + k) + a.amp * noise(x);
  if (a.sub) v = v - __ldg(a.sub + (size_t)c * a.nk + k);
  if (a.clamp0) v = fmax(v, 0.0);
  return (double)(T)v;
}

// Generic path (any slab size): one Philox group per thread iteration, per-element level lookup.
template <typename T>
__global__ void __launch_bounds__(256) les_state_generic_kernel(const K5Args a) {
  const long long total = a.ngrp * a.ncol;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t / a.ngrp);
    const long long g = t - (long long)c * a.ngrp;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(a.col0 + c), a.stream_id), a.seed,
                                  0x5BD1E995u);
    const long long e0 = g * 4;
    T* out = static_cast<T*>(a.vol) + (size_t)c * a.nE;
    const uint32_t x[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (e0 + i < a.nE) out[e0 + i] = (T)value<T>(a, c, e0 + i, x[i]);
  }
}

// Fast path (slab elements % 4 == 0, 16-byte aligned volume): blocks walk (column, level) slabs, so the
// level's profile values are loaded once per slab and no per-element division is needed; each thread
// produces one Philox group = one 16-byte (float32) or two 16-byte (float64) coalesced stores.
// The kernel is bound by instruction issue (Philox is ~10 integer instructions per element), so everything
// around the ten rounds is pared down: SUB / CLAMP are template flags, the 64-bit products are single
// IMAD.WIDE, and the noise n = 2*(x>>8)*2^-24 - 1 is formed without an int->double conversion:
// with m = x>>8, the double whose bits are 0x43300000:m equals 2^52 + m, so (that - (2^52 + 2^23)) = m - 2^23
// exactly and amp*n = (amp*2^-23)*(m - 2^23) is the same real product, hence the same rounding, as numpy's.
__device__ __forceinline__ uint4 philox4x32_10_wide(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)kM0 * c0, p1 = (uint64_t)kM1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += kW0;
    k1 += kW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

template <typename T, bool SUB, bool CLAMP>
__global__ void __launch_bounds__(256) les_state_slab_kernel(const K5Args a, const double amp23) {
  const long long nslab = (long long)a.ncol * a.nk;
  const int gps = (int)(a.S >> 2);  // Philox groups per slab
  constexpr double kMagic = 4503599627370496.0 + 8388608.0;  // 2^52 + 2^23
  for (long long sl = blockIdx.x; sl < nslab; sl += gridDim.x) {
    const int c = (int)(sl / a.nk), k = (int)(sl - (long long)c * a.nk);
    const double base = __ldg(a.prof + sl);
    const double sub = SUB ? __ldg(a.sub + sl) : 0.0;
    T* out = static_cast<T*>(a.vol) + (size_t)sl * a.S;
    const unsigned long long g0 = (unsigned long long)k * gps;  // first group of this slab within the column
    const uint32_t col = (uint32_t)(a.col0 + c);
    for (int gl = threadIdx.x; gl < gps; gl += blockDim.x) {
      const unsigned long long g = g0 + gl;
      const uint4 r = philox4x32_10_wide((uint32_t)g, (uint32_t)(g >> 32), col, a.stream_id, a.seed, 0x5BD1E995u);
      const uint32_t x[4] = {r.x, r.y, r.z, r.w};
      T v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double dm = __hiloint2double(0x43300000, (int)(x[i] >> 8)) - kMagic;  // (x>>8) - 2^23, exact
        double d = base + amp23 * dm;
        if constexpr (SUB) d = d - sub;
        if constexpr (CLAMP) d = fmax(d, 0.0);
        v[i] = (T)d;
      }
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(out + 4 * gl) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        *reinterpret_cast<double2*>(out + 4 * gl) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(out + 4 * gl + 2) = make_double2(v[2], v[3]);
      }
    }
  }
}

template <typename T>
void launch_slab(const K5Args& a, double amp23, int grid, cudaStream_t st) {
  const bool sub = a.sub != nullptr, clamp = a.clamp0 != 0;
  if (sub && clamp) les_state_slab_kernel<T, true, true><<<grid, 256, 0, st>>>(a, amp23);
  else if (sub) les_state_slab_kernel<T, true, false><<<grid, 256, 0, st>>>(a, amp23);
  else if (clamp) les_state_slab_kernel<T, false, true><<<grid, 256, 0, st>>>(a, amp23);
  else les_state_slab_kernel<T, false, false><<<grid, 256, 0, st>>>(a, amp23);
}

}  // namespace

extern "C" int spc_set_les_state(spc_handle h, const double* prof, double amp, uint32_t stream_id, uint32_t seed, int col0,
                                 const double* sub, int clamp0, void* vol, int dtype, int ncol, int nx, int ny, int nk,
                                 void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(dtype == SPC_F32 || dtype == SPC_F64, SPC_ERR_ARG, "spc_set_les_state: bad dtype %d", dtype);
  SPC_REQUIRE(ncol >= 0 && nx > 0 && ny > 0 && nk > 0, SPC_ERR_ARG, "spc_set_les_state: bad shape");
  if (ncol == 0) return SPC_OK;
  SPC_REQUIRE(prof && vol, SPC_ERR_ARG, "spc_set_les_state: NULL pointer");
  spc::DeviceGuard guard(h->device);
  K5Args a;
  a.prof = prof; a.sub = sub; a.vol = vol; a.amp = amp; a.stream_id = stream_id; a.seed = seed;
  a.col0 = col0; a.clamp0 = clamp0; a.ncol = ncol; a.nk = nk;
  a.S = (long long)nx * ny;
  a.nE = a.S * nk;
  a.ngrp = (a.nE + 3) / 4;
  // amp * 2^-23 must be exact for the slab kernel's noise form (it is unless amp is within 2^23 of the subnormals)
  const double amp23 = ldexp(amp, -23);
  const bool fast = (a.S % 4 == 0) && (reinterpret_cast<uintptr_t>(vol) % 16 == 0) && ldexp(amp23, 23) == amp;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (fast) {
    const int grid = (int)std::min<long long>((long long)ncol * nk, (long long)h->num_sms * 8);
    if (dtype == SPC_F32) launch_slab<float>(a, amp23, grid, st);
    else launch_slab<double>(a, amp23, grid, st);
  } else {
    const long long total = a.ngrp * ncol;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 16);
    if (dtype == SPC_F32) les_state_generic_kernel<float><<<grid, 256, 0, st>>>(a);
    else les_state_generic_kernel<double><<<grid, 256, 0, st>>>(a);
  }
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}
