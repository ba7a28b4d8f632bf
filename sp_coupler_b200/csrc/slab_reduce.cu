// K1 slab_reduce: horizontal slab means of the five LES volumes + above-threshold ql count/mask.
//
// Replaces the LES-side reductions the reference requests in spcpl.get_les_profiles
// (splib/spcpl.py:747-759) and les.get_cloudfraction (spcpl.py:28,765). HBM-bound: every volume
// byte is read exactly once; algorithmic bytes = 5*nx*ny*nk*sizeof(T) per column.
//
// Fast path (KJI layout, slab bytes % 16 == 0): persistent grid, one CTA per SM, every WARP owns a
// private ring of shared-memory stages fed by 1-D TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx). Lane 0 keeps the ring full; the warp reduces a landed chunk in 4 KB sub-blocks with
// conflict-free 128-bit LDS, accumulating in float64 (so the mean is order-insensitive to ~1e-16),
// and finishes a slab with a shuffle butterfly. No block-wide barriers, no atomics, fixed order.
//
// Kernels in this file
//   slab_reduce_tma_kernel<T, Ring>        KJI, slabs >= 1 KB and a multiple of 16 bytes (the production path)
//   slab_reduce_tma_pair_kernel<T, W>      KJI, 4 KB slabs: two slabs per 8 KB copy
//   slab_reduce_generic_kernel<T>          KJI, any other slab size / alignment (one warp per slab, scalar loads)
//   slab_reduce_ijk_tma_kernel<T, SLOTS>   IJK (k fastest), per-warp TMA rings over whole (field, column) items
//   slab_reduce_ijk_kernel<T, VEC>         IJK, shapes outside the TMA plan (one CTA per item)
#include <math.h>

#include <type_traits>

#include "spc_common.cuh"

namespace {

// Shape of the per-warp TMA ring: WARPS warps per CTA (one CTA per SM), STAGES stages of CHUNK bytes
// each (CHUNK = NSUB x 4 KB sub-blocks; a sub-block is what one warp reduces per pass and one cloud-mask
// word per lane). BLOCKED: a warp owns a contiguous run of slabs instead of every nwarps-th slab.
template <int WARPS_, int CHUNK_, int STAGES_, bool BLOCKED_ = false, int HINT_ = 1, int UNROLL_ = 1>
struct Ring {
  static constexpr int kWarps = WARPS_, kChunk = CHUNK_, kStages = STAGES_;
  static constexpr bool kBlocked = BLOCKED_;
  static constexpr int kHint = HINT_;      // L2 policy of the bulk copies: 0 evict_first, 1 evict_normal, 2 no cache hint, 3 evict_last
  static constexpr int kUnroll = UNROLL_;  // sub-blocks of a chunk reduced per loop trip
  static_assert(CHUNK_ % 4096 == 0, "chunk must be a multiple of the 4 KB sub-block");
  static_assert((size_t)WARPS_ * STAGES_ * (CHUNK_ + 8) <= 227 * 1024, "ring exceeds shared memory");
};
constexpr int kSubBytes = 4096;
constexpr uint32_t kFull = 0xffffffffu;

struct K1Args {
  const void *v0, *v1, *v2, *v3, *v4;  // THL,QT,QL,U,V (named members: no local-memory copy for vol[f])
  double* prof;
  int32_t* cnt;
  uint32_t* mask;
  double thr;
  float thr_f;          // smallest float32 x with (double)x > thr (NaN if none): x >= thr_f  <=>  (double)x > thr, exactly
  long long per_field;  // ncol*nk slabs per field
  long long total;      // 5*ncol*nk
  int S;                // elements per slab
  int slab_bytes;
  int nch;              // TMA chunks per slab
  int nsub;             // 4 KB sub-blocks per slab (= cloud-mask words per lane per slab)
};

__device__ __forceinline__ const void* field_ptr(const K1Args& a, int f) {
  return f == 0 ? a.v0 : f == 1 ? a.v1 : f == 2 ? a.v2 : f == 3 ? a.v3 : a.v4;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_evict_normal_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ __attribute__((unused)) void tma_bulk_g2s_nohint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// One lane's share of a landed 4 KB sub-block: 16-byte vectors lane, lane+32, ... of `nvec` (<= 256).
// Returns the lane's cloud bits (ql only): bit it*4+c (float32) / it*2+c (float64).
template <typename T, bool QL>
__device__ __forceinline__ uint32_t consume(const uint8_t* buf, int nvec, int lane, double (&acc)[4], int& cnt, double thr) {
  constexpr int kIter = kSubBytes / 16 / 32;  // 8
  constexpr int kFullVec = kSubBytes / 16;    // 256
  uint32_t bits = 0;
  if constexpr (sizeof(T) == 4) {
    const float4* b = reinterpret_cast<const float4*>(buf);
    float4 v[kIter];
    if (nvec == kFullVec) {
#pragma unroll
      for (int it = 0; it < kIter; ++it) v[it] = b[it * 32 + lane];
    } else {
#pragma unroll
      for (int it = 0; it < kIter; ++it)
        v[it] = (it * 32 + lane < nvec) ? b[it * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      double d0 = (double)v[it].x, d1 = (double)v[it].y, d2 = (double)v[it].z, d3 = (double)v[it].w;
      acc[0] += d0;
      acc[1] += d1;
      acc[2] += d2;
      acc[3] += d3;
      if constexpr (QL) {
        // padding lanes hold 0.0 and must not count when thr < 0
        bool in = (nvec == kFullVec) || (it * 32 + lane < nvec);
        uint32_t m = (uint32_t)(in && d0 > thr) | ((uint32_t)(in && d1 > thr) << 1) |
                     ((uint32_t)(in && d2 > thr) << 2) | ((uint32_t)(in && d3 > thr) << 3);
        cnt += __popc(m);
        bits |= m << (it * 4);
      }
    }
  } else {
    const double2* b = reinterpret_cast<const double2*>(buf);
    double2 v[kIter];
    if (nvec == kFullVec) {
#pragma unroll
      for (int it = 0; it < kIter; ++it) v[it] = b[it * 32 + lane];
    } else {
#pragma unroll
      for (int it = 0; it < kIter; ++it) v[it] = (it * 32 + lane < nvec) ? b[it * 32 + lane] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      acc[(2 * it) & 3] += v[it].x;
      acc[(2 * it + 1) & 3] += v[it].y;
      if constexpr (QL) {
        bool in = (nvec == kFullVec) || (it * 32 + lane < nvec);
        uint32_t m = (uint32_t)(in && v[it].x > thr) | ((uint32_t)(in && v[it].y > thr) << 1);
        cnt += __popc(m);
        bits |= m << (it * 2);
      }
    }
  }
  return bits;
}

template <typename T, typename R>
__global__ void __launch_bounds__(R::kWarps * 32, 1) slab_reduce_tma_kernel(const K1Args a) {
  constexpr int kWarps = R::kWarps, kChunk = R::kChunk, kStages = R::kStages;
  extern __shared__ __align__(128) uint8_t smem[];
  // [kWarps][kStages][kChunk] data, then [kWarps][kStages] mbarriers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWarps * kStages * kChunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wbuf = smem + (size_t)warp * kStages * kChunk;
  const uint32_t wbuf_s = smem_u32(wbuf);
  const uint32_t wbar_s = smem_u32(bars + warp * kStages);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(wbar_s + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  // slabs of this warp: g(i) = g0 + i*gstride, i < nslab
  const long long nwarps = (long long)gridDim.x * kWarps;
  const long long gw = (long long)blockIdx.x * kWarps + warp;
  long long g0, gstride, nslab;
  if constexpr (R::kBlocked) {
    const long long base = a.total / nwarps, rem = a.total % nwarps;
    g0 = gw * base + min(gw, rem);
    nslab = base + (gw < rem ? 1 : 0);
    gstride = 1;
  } else {
    g0 = gw;
    gstride = nwarps;
    nslab = gw < a.total ? (a.total - gw + nwarps - 1) / nwarps : 0;
  }
  if (nslab == 0) return;
  const long long nq = nslab * a.nch;  // chunks this warp streams
  const uint64_t pol = R::kHint == 0 ? l2_evict_first_policy() : R::kHint == 3 ? l2_evict_last_policy() : l2_evict_normal_policy();

  // (field, slab-within-field) of a slab index advance incrementally with the stride: no 64-bit division per slab
  const int f0 = (int)(g0 / a.per_field);
  const long long rem0 = g0 - (long long)f0 * a.per_field;
  // producer state (lane 0): next chunk to issue
  int pf = f0;            // field of the slab being fetched
  long long prem = rem0;  // its index within the field (= c*nk + k)
  int pj = 0;             // chunk within slab
  auto issue = [&](int stage) {
    const uint8_t* src = static_cast<const uint8_t*>(field_ptr(a, pf)) + (size_t)prem * a.slab_bytes + (size_t)pj * kChunk;
    const int bytes = min(kChunk, a.slab_bytes - pj * kChunk);
    const uint32_t bar = wbar_s + 8 * stage;
    mbar_arrive_expect_tx(bar, (uint32_t)bytes);
    if constexpr (R::kHint == 2) tma_bulk_g2s_nohint(wbuf_s + stage * kChunk, src, (uint32_t)bytes, bar);
    else tma_bulk_g2s(wbuf_s + stage * kChunk, src, (uint32_t)bytes, bar, pol);
    if (++pj == a.nch) {
      pj = 0;
      prem += gstride;
      while (prem >= a.per_field) {
        prem -= a.per_field;
        ++pf;
      }
    }
  };
  if (lane == 0) {
    const int pre = (int)min((long long)kStages, nq);
    for (int s = 0; s < pre; ++s) issue(s);
  }

  int stage = 0;
  uint32_t parity = 0;
  long long q = 0;
  int f = f0;
  long long rem = rem0;  // = c*nk + k
  for (long long i = 0; i < nslab; ++i) {
    const long long g = g0 + i * gstride;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int cnt = 0;
    const bool is_ql = (f == SPC_QL) && (a.cnt != nullptr || a.mask != nullptr);
    uint32_t* mrow = a.mask ? a.mask + (size_t)rem * a.nsub * 32 + lane : nullptr;
    int sub_id = 0;  // 4 KB sub-block ordinal within the slab
    for (int j = 0; j < a.nch; ++j, ++q) {
      const uint32_t bar = wbar_s + 8 * stage;
      while (!mbar_try_wait(bar, parity)) {
      }
      const int bytes = min(kChunk, a.slab_bytes - j * kChunk);
      const uint8_t* buf = wbuf + stage * kChunk;
#pragma unroll R::kUnroll
      for (int off = 0; off < bytes; off += kSubBytes, ++sub_id) {
        const int nvec = min(kSubBytes, bytes - off) >> 4;
        if (is_ql) {
          const uint32_t bits = consume<T, true>(buf + off, nvec, lane, acc, cnt, a.thr);
          if (mrow) mrow[(size_t)sub_id * 32] = bits;
        } else {
          consume<T, false>(buf + off, nvec, lane, acc, cnt, a.thr);
        }
      }
      __syncwarp();  // every lane is done reading this stage before it is refilled
      if (lane == 0 && q + kStages < nq) issue(stage);
      if (++stage == kStages) {
        stage = 0;
        parity ^= 1;
      }
    }
    double s = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
    if (is_ql && a.cnt) {
      int c = warp_sum(cnt);
      if (lane == 0) a.cnt[rem] = c;
    }
    if (lane == 0) a.prof[g] = s / (double)a.S;
    rem += gstride;
    while (rem >= a.per_field) {
      rem -= a.per_field;
      ++f;
    }
  }
}

// 4 KB slabs (32 x 32 float32): a TMA copy of one slab is only half the size the copy engine likes, and twice as
// many warps are needed to keep the same bytes in flight. This variant streams PAIRS of consecutive slabs as one
// 8 KB copy per warp (12 warps), reduces the two halves into separate accumulators, hands the stage back and only
// then runs the two slab epilogues (shuffle butterflies, divide, stores), so they overlap with the next fetch.
template <typename T, int WARPS, int STAGES = 1>
__global__ void __launch_bounds__(WARPS * 32, 1) slab_reduce_tma_pair_kernel(const K1Args a) {
  constexpr int kChunk = 2 * kSubBytes;
  extern __shared__ __align__(128) uint8_t smem[];
  // [WARPS][STAGES][kChunk] data, then [WARPS][STAGES] mbarriers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * STAGES * kChunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wbuf = smem + (size_t)warp * STAGES * kChunk;
  const uint32_t wbuf_s = smem_u32(wbuf);
  const uint32_t bar0 = smem_u32(bars + warp * STAGES);
  if (lane == 0) {
#pragma unroll
    for (int sg = 0; sg < STAGES; ++sg) mbar_init(bar0 + 8 * sg, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  // pair p = slabs 2p, 2p+1 (same field: per_field is even on this path); warp w of the grid takes pairs w, w+nwarps, ...
  const long long nwarps = (long long)gridDim.x * WARPS, npairs = a.total >> 1, half_field = a.per_field >> 1;
  const long long p0 = (long long)blockIdx.x * WARPS + warp;
  if (p0 >= npairs) return;
  const long long npair = (npairs - p0 + nwarps - 1) / nwarps;
  const uint64_t pol = l2_evict_normal_policy();
  int f = (int)(p0 / half_field);               // consumer cursor (every lane carries it): field, pair within the field
  long long rem = p0 - (long long)f * half_field;
  int pf = f;                                   // producer cursor (lane 0 advances it when it issues a copy)
  long long prem = rem;
  auto issue = [&](int sg) {
    const uint8_t* src = static_cast<const uint8_t*>(field_ptr(a, pf)) + (size_t)prem * kChunk;
    const uint32_t bar = bar0 + 8 * sg;
    mbar_arrive_expect_tx(bar, (uint32_t)kChunk);
    tma_bulk_g2s(wbuf_s + sg * kChunk, src, (uint32_t)kChunk, bar, pol);
    prem += nwarps;
    while (prem >= half_field) {
      prem -= half_field;
      ++pf;
    }
  };
  if (lane == 0) {
#pragma unroll
    for (int sg = 0; sg < STAGES; ++sg)
      if (sg < npair) issue(sg);
  }
  uint32_t parity = 0;
  int stage = 0;
  for (long long i = 0; i < npair; ++i) {
    const long long s0 = 2 * rem;               // first slab of the pair within the field (= c*nk + k)
    const bool is_ql = (f == SPC_QL) && (a.cnt != nullptr || a.mask != nullptr);
    double accA[4] = {0.0, 0.0, 0.0, 0.0}, accB[4] = {0.0, 0.0, 0.0, 0.0};
    int cntA = 0, cntB = 0;
    while (!mbar_try_wait(bar0 + 8 * stage, parity)) {
    }
    const uint8_t* sbuf = wbuf + stage * kChunk;
    if (is_ql) {
      const uint32_t bitsA = consume<T, true>(sbuf, kSubBytes >> 4, lane, accA, cntA, a.thr);
      const uint32_t bitsB = consume<T, true>(sbuf + kSubBytes, kSubBytes >> 4, lane, accB, cntB, a.thr);
      if (a.mask) {
        a.mask[(size_t)s0 * 32 + lane] = bitsA;
        a.mask[(size_t)(s0 + 1) * 32 + lane] = bitsB;
      }
    } else {
      consume<T, false>(sbuf, kSubBytes >> 4, lane, accA, cntA, a.thr);
      consume<T, false>(sbuf + kSubBytes, kSubBytes >> 4, lane, accB, cntB, a.thr);
    }
    __syncwarp();  // every lane is done reading the stage before it is refilled
    if (lane == 0 && i + STAGES < npair) issue(stage);
    if (++stage == STAGES) {
      stage = 0;
      parity ^= 1;
    }
    const double sA = warp_sum((accA[0] + accA[1]) + (accA[2] + accA[3]));
    const double sB = warp_sum((accB[0] + accB[1]) + (accB[2] + accB[3]));
    if (is_ql && a.cnt) {
      const int cA = warp_sum(cntA), cB = warp_sum(cntB);
      if (lane == 0) {
        a.cnt[s0] = cA;
        a.cnt[s0 + 1] = cB;
      }
    }
    if (lane == 0) {
      const long long g = (long long)f * a.per_field + s0;
      a.prof[g] = sA / (double)a.S;
      a.prof[g + 1] = sB / (double)a.S;
    }
    rem += nwarps;
    while (rem >= half_field) {
      rem -= half_field;
      ++f;
    }
  }
}

// Generic KJI path: any slab size / alignment. One warp per slab, scalar read-only loads.
template <typename T>
__global__ void __launch_bounds__(256) slab_reduce_generic_kernel(const K1Args a) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int mw = (a.S + 31) >> 5;
  for (long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < a.total; g += nwarps) {
    const int f = (int)(g / a.per_field);
    const long long rem = g - (long long)f * a.per_field;
    const T* p = static_cast<const T*>(field_ptr(a, f)) + (size_t)rem * a.S;
    const bool is_ql = (f == SPC_QL) && (a.cnt != nullptr || a.mask != nullptr);
    double acc = 0.0;
    int cnt = 0;
    for (int base = 0; base < a.S; base += 32) {
      const int e = base + lane;
      const bool in = e < a.S;
      const double v = in ? (double)__ldg(p + e) : 0.0;
      acc += v;
      if (is_ql) {
        const uint32_t w = __ballot_sync(kFull, in && v > a.thr);
        cnt += __popc(w);
        if (a.mask && lane == 0) a.mask[(size_t)rem * mw + (base >> 5)] = w;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      a.prof[g] = acc / (double)a.S;
      if (is_ql && a.cnt) a.cnt[rem] = cnt;
    }
  }
}

// IJK layout ([ncol][nx][ny][nk], k fastest): one CTA per (field, column); thread (r, kk) walks the
// horizontal points r, r+R, ... of its level(s) kk with coalesced loads along k, then the R partial
// sums per level are combined through shared memory in a fixed order.
constexpr int kMaskRows = 4;  // row groups per barrier pair in the masked IJK loop

template <typename T, int VEC>
__global__ void __launch_bounds__(1024) slab_reduce_ijk_kernel(const K1Args a, int nk, int ncol, int nkv, int R) {
  extern __shared__ __align__(128) uint8_t smem[];
  double* ssum = reinterpret_cast<double*>(smem);                  // [R][nk]
  int* scnt = reinterpret_cast<int*>(ssum + (size_t)R * nk);       // [R][nk]
  const int kw = (nk + 31) >> 5;                                   // mask words per horizontal point
  uint32_t* rowbits = reinterpret_cast<uint32_t*>(scnt + (size_t)R * nk);  // [kMaskRows*R][kw]
  const int f = blockIdx.x / ncol, c = blockIdx.x - f * ncol;
  const int kk = threadIdx.x % nkv, r = threadIdx.x / nkv;
  const bool is_ql = (f == SPC_QL) && (a.cnt != nullptr || a.mask != nullptr);
  const bool do_mask = (f == SPC_QL) && a.mask != nullptr;         // block-uniform
  const T* p = static_cast<const T*>(field_ptr(a, f)) + (size_t)c * a.S * nk;
  // IJK cloud mask: [ncol][S][kw] words, bit k%32 of word k/32 of horizontal point `row`
  uint32_t* mcol = do_mask ? a.mask + (size_t)c * a.S * kw : nullptr;
  double acc[VEC];
  int cnt[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    acc[v] = 0.0;
    cnt[v] = 0;
  }
  if (do_mask) {
    for (int i = threadIdx.x; i < R * kMaskRows * kw; i += blockDim.x) rowbits[i] = 0u;
    __syncthreads();
  }
  // one horizontal point: accumulate this thread's VEC levels, return their cloud flags
  auto point = [&](int row) -> uint32_t {
    uint32_t flags = 0;
    if constexpr (VEC == 1) {
      const double d = (double)__ldg(p + (size_t)row * nk + kk);
      acc[0] += d;
      flags = (d > a.thr);
    } else if constexpr (sizeof(T) == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p + (size_t)row * nk) + kk);
      const double d[4] = {(double)q.x, (double)q.y, (double)q.z, (double)q.w};
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        acc[v] += d[v];
        flags |= (uint32_t)(d[v] > a.thr) << v;
      }
    } else {
      const double2 q = __ldg(reinterpret_cast<const double2*>(p + (size_t)row * nk) + kk);
      acc[0] += q.x;
      acc[1] += q.y;
      flags = (uint32_t)(q.x > a.thr) | ((uint32_t)(q.y > a.thr) << 1);
    }
    if (is_ql) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) cnt[v] += (flags >> v) & 1u;
    }
    return flags;
  };
  if (!do_mask) {
    // streaming loop without block barriers: the compiler keeps several rows of loads in flight
    if (r < R) {
#pragma unroll 8
      for (int row = r; row < a.S; row += R) point(row);
    }
  } else {
    // kMaskRows row groups per barrier pair: loads of 4 rows in flight, 4x fewer block barriers
    const int step = R * kMaskRows;
    for (int row0 = 0; row0 < a.S; row0 += step) {                 // block-uniform trip count
      uint32_t flags[kMaskRows];
#pragma unroll
      for (int u = 0; u < kMaskRows; ++u) {
        const int row = row0 + u * R + r;
        flags[u] = (r < R && row < a.S) ? point(row) : 0u;
      }
      const int k0 = kk * VEC;                                     // VEC divides 32: no word straddling
#pragma unroll
      for (int u = 0; u < kMaskRows; ++u)
        if (flags[u]) atomicOr(&rowbits[(u * R + r) * kw + (k0 >> 5)], flags[u] << (k0 & 31));
      __syncthreads();
      for (int i = threadIdx.x; i < step * kw; i += blockDim.x) {
        const int rr = i / kw;
        if (row0 + rr < a.S) mcol[(size_t)(row0 + rr) * kw + (i - rr * kw)] = rowbits[i];
        rowbits[i] = 0u;
      }
      __syncthreads();
    }
  }
  if (r < R) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      ssum[(size_t)r * nk + kk * VEC + v] = acc[v];
      scnt[(size_t)r * nk + kk * VEC + v] = cnt[v];
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    double s = 0.0;
    int n = 0;
    for (int rr = 0; rr < R; ++rr) {
      s += ssum[(size_t)rr * nk + k];
      n += scnt[(size_t)rr * nk + k];
    }
    a.prof[((size_t)f * ncol + c) * nk + k] = s / (double)a.S;
    if (is_ql && a.cnt) a.cnt[(size_t)c * nk + k] = n;
  }
}

// IJK fast path: per-warp TMA rings over the contiguous [S][nk] block of one (field, column) item.
// A 16-byte vector v of the block holds levels (v % nkv)*VEC .. +VEC-1, so with a period of
// P = lcm(nkv, 32) vectors lane `l` sees the same level group in slot `it` of every period:
// (it*32 + l) % nkv. Each lane therefore keeps SLOTS x VEC float64 partial sums in registers and never
// needs to know which horizontal point it is looking at. Persistent CTAs own whole items; the chunks of
// a CTA's items form one sequence that is dealt round-robin to its warps, so the rings never drain at an
// item boundary. When a warp has finished its share of an item it parks its partials in shared memory
// and carries on with the next item; once all warps have parked, each warp combines its slice of the
// levels in a fixed order (bit-reproducible, no atomics on data) the next time it looks. No block-wide
// barrier in the steady state. Items are numbered column-major (item = c*5 + f) so that at any moment the
// SMs are spread over all five fields and the heavier ql items (count + mask) overlap with the others.
template <int WARPS_, int CHUNK_, int BATCH5_ = 2, int STAGES_ = 1>
struct IjkRing {
  static constexpr int kWarps = WARPS_, kChunk = CHUNK_, kBatch5 = BATCH5_;  // kBatch5: periods per pass when SLOTS >= 4
  static constexpr int kStages = STAGES_;   // ring stages per warp: > 1 keeps a copy in flight while the warp reduces a chunk
};

struct IjkArgs {
  int nk, ncol, nkv;          // levels, columns, 16-byte vectors per horizontal point
  int chunk_bytes;            // whole periods per TMA chunk, <= kChunk
  int nch;                    // chunks per item
  long long item_bytes;       // S*nk*sizeof(T)
  long long nitems;           // 5*ncol
  int want_mask;
};

// UU periods of a landed chunk: all 128-bit LDS are issued before the first conversion.
template <typename T, int SLOTS, int UU, bool QL, bool MASK>
__device__ __forceinline__ void ijk_periods(const uint8_t* pb, int lane, double thr, uint32_t* mw,
                                            double (&acc)[SLOTS][16 / sizeof(T)], int (&cnt)[SLOTS][16 / sizeof(T)]) {
  constexpr int VEC = 16 / (int)sizeof(T), LPW = 32 / VEC, P = SLOTS * 32;
  if constexpr (sizeof(T) == 4) {
    float4 v[UU][SLOTS];
#pragma unroll
    for (int u = 0; u < UU; ++u)
#pragma unroll
      for (int it = 0; it < SLOTS; ++it) v[u][it] = reinterpret_cast<const float4*>(pb)[u * P + it * 32 + lane];
#pragma unroll
    for (int u = 0; u < UU; ++u)
#pragma unroll
      for (int it = 0; it < SLOTS; ++it) {
        const double d0 = (double)v[u][it].x, d1 = (double)v[u][it].y, d2 = (double)v[u][it].z, d3 = (double)v[u][it].w;
        acc[it][0] += d0;
        acc[it][1] += d1;
        acc[it][2] += d2;
        acc[it][3] += d3;
        if constexpr (QL) {
          const uint32_t b0 = d0 > thr, b1 = d1 > thr, b2 = d2 > thr, b3 = d3 > thr;
          cnt[it][0] += b0;
          cnt[it][1] += b1;
          cnt[it][2] += b2;
          cnt[it][3] += b3;
          if constexpr (MASK) {
            uint32_t word = (b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)) << (4 * (lane & 7));
            word |= __shfl_xor_sync(kFull, word, 1);  // (redux.sync.or over the 8-lane group measured 16 % slower)
            word |= __shfl_xor_sync(kFull, word, 2);
            word |= __shfl_xor_sync(kFull, word, 4);
            if ((lane & 7) == 0) mw[(u * P + it * 32) / LPW] = word;
          }
        }
      }
  } else {
    double2 v[UU][SLOTS];
#pragma unroll
    for (int u = 0; u < UU; ++u)
#pragma unroll
      for (int it = 0; it < SLOTS; ++it) v[u][it] = reinterpret_cast<const double2*>(pb)[u * P + it * 32 + lane];
#pragma unroll
    for (int u = 0; u < UU; ++u)
#pragma unroll
      for (int it = 0; it < SLOTS; ++it) {
        acc[it][0] += v[u][it].x;
        acc[it][1] += v[u][it].y;
        if constexpr (QL) {
          const uint32_t b0 = v[u][it].x > thr, b1 = v[u][it].y > thr;
          cnt[it][0] += b0;
          cnt[it][1] += b1;
          if constexpr (MASK) {
            uint32_t word = (b0 | (b1 << 1)) << (2 * (lane & 15));
            word |= __shfl_xor_sync(kFull, word, 1);
            word |= __shfl_xor_sync(kFull, word, 2);
            word |= __shfl_xor_sync(kFull, word, 4);
            word |= __shfl_xor_sync(kFull, word, 8);
            if ((lane & 15) == 0) mw[(u * P + it * 32) / LPW] = word;
          }
        }
      }
  }
}

// Transpose of an NL x NL matrix of EB-bit elements held one row per lane by NL consecutive lanes (NL * EB == 32):
// afterwards element i of lane j is what element j of lane i was. log2(NL) butterfly stages of one shuffle each.
template <int NL, int EB>
__device__ __forceinline__ uint32_t lane_transpose_bits(uint32_t x, int lane) {
  static_assert(NL * EB == 32, "one 32-bit word per lane");
#pragma unroll
  for (int s = 0; (1 << s) < NL; ++s) {
    const int d = 1 << s, sh = EB << s;
    // elements whose index has bit s clear: runs of sh ones alternating with runs of sh zeros
    const uint32_t m = sh == 2 ? 0x33333333u : sh == 4 ? 0x0F0F0F0Fu : sh == 8 ? 0x00FF00FFu : 0x0000FFFFu;
    const uint32_t y = __shfl_xor_sync(kFull, x, d);
    x = (lane & d) ? (((y >> sh) & m) | (x & ~m)) : ((x & m) | ((y & m) << sh));
  }
  return x;
}

// The ql chunk with the per-point cloud mask. A mask word (32 levels of one horizontal point) is spread over the NL =
// 32 / VEC lanes that hold the point's vectors in ONE load, so assembling it per load costs a log2(NL)-step shuffle
// butterfly for every vector (the round-1 form: 3 dependent shuffles per float4). Here every lane instead collects
// its own VEC flag bits of up to NL consecutive loads in a private word (compile-time bit positions), and one NL x NL
// bit-block transpose across the lane group turns the private words into the natural mask words of those loads:
// log2(NL) shuffles per group of loads, and the lanes then store consecutive words. Same mask bits, same layout.
// The groups are cut at compile time inside a batch of UU periods (UU * SLOTS loads): full groups of NL, then the rest.
template <typename T, int SLOTS, int UU>
__device__ __forceinline__ void ijk_periods_mask(const uint8_t* pb, int lane, double thr, float thr_f, uint32_t* mw0,
                                                 double (&acc)[SLOTS][16 / sizeof(T)], int (&cnt)[SLOTS][16 / sizeof(T)]) {
  constexpr int VEC = 16 / (int)sizeof(T), NL = 32 / VEC, P = SLOTS * 32, NLOADS = UU * SLOTS;
  using V = std::conditional_t<sizeof(T) == 4, float4, double2>;
  V v[UU][SLOTS];
#pragma unroll
  for (int u = 0; u < UU; ++u)
#pragma unroll
    for (int it = 0; it < SLOTS; ++it) v[u][it] = reinterpret_cast<const V*>(pb)[u * P + it * 32 + lane];
  uint32_t W = 0u;
#pragma unroll
  for (int t = 0; t < NLOADS; ++t) {           // load t of the batch = period t / SLOTS, slot t % SLOTS
    const int u = t / SLOTS, it = t % SLOTS;   // compile-time after unrolling
    uint32_t nib;
    if constexpr (sizeof(T) == 4) {
      const double d0 = (double)v[u][it].x, d1 = (double)v[u][it].y, d2 = (double)v[u][it].z, d3 = (double)v[u][it].w;
      acc[it][0] += d0;
      acc[it][1] += d1;
      acc[it][2] += d2;
      acc[it][3] += d3;
      const uint32_t b0 = v[u][it].x >= thr_f, b1 = v[u][it].y >= thr_f, b2 = v[u][it].z >= thr_f, b3 = v[u][it].w >= thr_f;
      cnt[it][0] += b0;
      cnt[it][1] += b1;
      cnt[it][2] += b2;
      cnt[it][3] += b3;
      nib = b0 | (b1 << 1) | (b2 << 2) | (b3 << 3);
    } else {
      acc[it][0] += v[u][it].x;
      acc[it][1] += v[u][it].y;
      const uint32_t b0 = v[u][it].x > thr, b1 = v[u][it].y > thr;
      cnt[it][0] += b0;
      cnt[it][1] += b1;
      nib = b0 | (b1 << 1);
    }
    W |= nib << (VEC * (t % NL));
    if ((t % NL) == NL - 1 || t == NLOADS - 1) {          // compile-time: a group of (t % NL) + 1 loads is complete
      const int t0 = t - (t % NL), n = (t % NL) + 1;
      const uint32_t x = lane_transpose_bits<NL, VEC>(W, lane);
      if (n == NL || (lane & (NL - 1)) < n) mw0[(t0 + (lane & (NL - 1))) * VEC + lane / NL] = x;   // VEC = 32 / NL words per load
      W = 0u;
    }
  }
}

template <typename T, int SLOTS, int BATCH5>
__device__ __forceinline__ void ijk_chunk_mask(const uint8_t* buf, int nper, int lane, double thr, float thr_f, uint32_t* mw0,
                                               double (&acc)[SLOTS][16 / sizeof(T)], int (&cnt)[SLOTS][16 / sizeof(T)]) {
  constexpr int VEC = 16 / (int)sizeof(T), P = SLOTS * 32;
  constexpr int kBatch = SLOTS >= 4 ? BATCH5 : 8 / SLOTS;
  int p = 0;                     // period cursor; a period is SLOTS loads = SLOTS * VEC mask words
#pragma unroll 1
  for (; p + kBatch <= nper; p += kBatch)
    ijk_periods_mask<T, SLOTS, kBatch>(buf + (size_t)p * (P * 16), lane, thr, thr_f, mw0 + (size_t)p * SLOTS * VEC, acc, cnt);
#pragma unroll 1
  for (; p < nper; ++p)
    ijk_periods_mask<T, SLOTS, 1>(buf + (size_t)p * (P * 16), lane, thr, thr_f, mw0 + (size_t)p * SLOTS * VEC, acc, cnt);
}

template <typename T, int SLOTS, int BATCH5, bool QL, bool MASK>
__device__ __forceinline__ void ijk_chunk(const uint8_t* buf, int nper, int lane, double thr, uint32_t* mw,
                                          double (&acc)[SLOTS][16 / sizeof(T)], int (&cnt)[SLOTS][16 / sizeof(T)]) {
  constexpr int VEC = 16 / (int)sizeof(T), LPW = 32 / VEC, P = SLOTS * 32;
  constexpr int kBatch = SLOTS >= 4 ? BATCH5 : 8 / SLOTS;
  int p = 0;
#pragma unroll 1
  for (; p + kBatch <= nper; p += kBatch)
    ijk_periods<T, SLOTS, kBatch, QL, MASK>(buf + (size_t)p * (P * 16), lane, thr, MASK ? mw + (p * P) / LPW : nullptr, acc, cnt);
#pragma unroll 1
  for (; p < nper; ++p)
    ijk_periods<T, SLOTS, 1, QL, MASK>(buf + (size_t)p * (P * 16), lane, thr, MASK ? mw + (p * P) / LPW : nullptr, acc, cnt);
}

template <typename T, int SLOTS, typename R>
__global__ void __launch_bounds__(R::kWarps * 32, 1) slab_reduce_ijk_tma_kernel(const K1Args a, const IjkArgs g) {
  constexpr int kWarps = R::kWarps, kChunk = R::kChunk, kStages = R::kStages;
  constexpr int VEC = 16 / (int)sizeof(T);     // levels per vector
  constexpr int LPW = 32 / VEC;                // lanes that share one 32-level mask word
  constexpr int P = SLOTS * 32;                // vectors per period
  extern __shared__ __align__(128) uint8_t smem[];
  // [kWarps][kStages][kChunk] ring | [kWarps][P][VEC] double partial sums | [kWarps][P][VEC] int partial counts |
  // [kWarps][kStages] mbarriers | flags
  double* psum = reinterpret_cast<double*>(smem + (size_t)kWarps * kStages * kChunk);
  int* pcnt = reinterpret_cast<int*>(psum + (size_t)kWarps * P * VEC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(pcnt + (size_t)kWarps * P * VEC);
  unsigned int* sync = reinterpret_cast<unsigned int*>(bars + kWarps * kStages);   // [0] partials parked, [1] level slices combined
  volatile unsigned int* vsync = sync;
  // the shuffle tells the compiler that the warp index is warp-uniform, so every branch below that depends on the
  // warp's item / chunk cursor is uniform too and the mask shuffles need no re-convergence wrappers
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint8_t* wbuf = smem + (size_t)warp * kStages * kChunk;
  const uint32_t wbuf_s = smem_u32(wbuf);
  const uint32_t bar0 = smem_u32(bars + warp * kStages);

  if (threadIdx.x == 0) {
    sync[0] = 0u;
    sync[1] = 0u;
  }
  if (lane == 0) {
#pragma unroll
    for (int sgi = 0; sgi < kStages; ++sgi) mbar_init(bar0 + 8 * sgi, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // items of this CTA: blockIdx.x + io*gridDim.x; its chunk sequence Q = io*nch + j; warp w takes Q = w (mod kWarps).
  // (io, j) are carried incrementally: no 64-bit division on the per-chunk path.
  const int my_items = blockIdx.x < g.nitems ? (int)((g.nitems - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  if (my_items == 0) return;
  const uint64_t pol = l2_evict_normal_policy();
  const int nch_mod = g.nch % kWarps;

  int pio = 0, pj = warp;  // producer cursor (lane 0): next chunk of this warp to fetch
  while (pj >= g.nch && pio < my_items) {
    pj -= g.nch;
    ++pio;
  }
  auto issue = [&](int stage) {
    const unsigned int item = blockIdx.x + (unsigned int)pio * gridDim.x;
    const unsigned int c = item / SPC_NFIELDS;
    const int f = (int)(item - c * SPC_NFIELDS);
    const uint8_t* src = static_cast<const uint8_t*>(field_ptr(a, f)) + (size_t)c * g.item_bytes + (size_t)pj * g.chunk_bytes;
    const uint32_t bytes = (uint32_t)min((long long)g.chunk_bytes, g.item_bytes - (long long)pj * g.chunk_bytes);
    const uint32_t bar = bar0 + 8 * stage;
    mbar_arrive_expect_tx(bar, bytes);
    tma_bulk_g2s(wbuf_s + stage * kChunk, src, bytes, bar, pol);
    pj += kWarps;
    while (pj >= g.nch && pio < my_items) {
      pj -= g.nch;
      ++pio;
    }
  };
  if (lane == 0) {
#pragma unroll
    for (int sgi = 0; sgi < kStages; ++sgi)
      if (pio < my_items) issue(sgi);
  }

  double acc[SLOTS][VEC];
  int cnt[SLOTS][VEC];
#pragma unroll
  for (int it = 0; it < SLOTS; ++it)
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      acc[it][c] = 0.0;
      cnt[it][c] = 0;
    }

  // Combines this warp's slice of the levels of item ordinal `pending` once every warp has parked its partials
  // (must = wait for that). Two lanes per level: lane halves sum the chunk classes [0, kWarps/2) and
  // [kWarps/2, kWarps) in order (within a class the period vectors of the level group), then lower + upper.
  int pending = -1;
  const int k_lo = (int)((long long)warp * g.nk / kWarps), k_hi = (int)((long long)(warp + 1) * g.nk / kWarps);
  auto combine = [&](bool must) {
    if (pending < 0) return;
    const unsigned int need = (unsigned int)kWarps * (unsigned int)(pending + 1);
    unsigned int have = 0;
    if (lane == 0) {
      have = vsync[0];
      while (must && have < need) have = vsync[0];
    }
    have = __shfl_sync(kFull, have, 0);
    if (have < need) return;
    __threadfence_block();
    const unsigned int item = blockIdx.x + (unsigned int)pending * gridDim.x;
    const unsigned int c = item / SPC_NFIELDS;
    const int f = (int)(item - c * SPC_NFIELDS);
    const bool is_ql = (f == SPC_QL) && a.cnt != nullptr;
    const int dup = P / g.nkv;  // period vectors holding the same level group
    const int half = lane >> 4, w0 = half * (kWarps / 2), w1 = half ? kWarps : kWarps / 2;
    for (int kb = k_lo; kb < k_hi; kb += 16) {       // warp-uniform trip count
      const int k = kb + (lane & 15);
      const bool live = k < k_hi;
      const int kg = k / VEC, cc = k - kg * VEC;
      double s = 0.0;
      int n = 0;
      if (live) {
        for (int w = w0; w < w1; ++w) {
          double t = 0.0;
          for (int d = 0; d < dup; ++d) {
            const size_t e = ((size_t)w * P + kg + (size_t)d * g.nkv) * VEC + cc;
            t += psum[e];
            if (is_ql) n += pcnt[e];
          }
          s += t;
        }
      }
      const double s_hi = __shfl_down_sync(kFull, s, 16);
      const int n_hi = __shfl_down_sync(kFull, n, 16);
      if (live && half == 0) {
        a.prof[((size_t)f * g.ncol + c) * g.nk + k] = (s + s_hi) / (double)a.S;
        if (is_ql) a.cnt[(size_t)c * g.nk + k] = n + n_hi;
      }
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicAdd(&sync[1], 1u);
    pending = -1;
  };

  // parks this warp's partials of item ordinal `io` under chunk class `cls`
  auto flush = [&](int io, int cls) {
    combine(true);                                                  // my slice of the previous item
    if (lane == 0)
      while (vsync[1] < (unsigned int)kWarps * (unsigned int)io) {  // every slice of it: the parking area is free
      }
    __syncwarp();
    double* ps = psum + ((size_t)cls * P) * VEC;
    int* pc = pcnt + ((size_t)cls * P) * VEC;
#pragma unroll
    for (int it = 0; it < SLOTS; ++it)
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        ps[(size_t)(it * 32 + lane) * VEC + c] = acc[it][c];
        pc[(size_t)(it * 32 + lane) * VEC + c] = cnt[it][c];
        acc[it][c] = 0.0;
        cnt[it][c] = 0;
      }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicAdd(&sync[0], 1u);
    pending = io;
  };

  uint32_t parity = 0;
  int stage = 0;
  // consumer cursor: chunk j of item ordinal io; cls = the chunk class (j % kWarps) this warp serves in item io -
  // a rotation of the warp index, so the combine order, hence every bit of the result, does not depend on where
  // the item sits in the batch
  int io = 0, j = warp, cls = warp;
  for (;;) {
    while (j >= g.nch && io < my_items) {  // this warp is done with item io
      flush(io, cls);
      j -= g.nch;
      ++io;
      cls = (cls - nch_mod + kWarps) % kWarps;
    }
    if (io >= my_items) break;
    combine(false);
    const unsigned int item = blockIdx.x + (unsigned int)io * gridDim.x;
    const unsigned int c = item / SPC_NFIELDS;
    const int f = (int)(item - c * SPC_NFIELDS);
    const bool is_ql = (f == SPC_QL) && (a.cnt != nullptr || a.mask != nullptr);
    const bool do_mask = (f == SPC_QL) && g.want_mask;
    const int bytes = (int)min((long long)g.chunk_bytes, g.item_bytes - (long long)j * g.chunk_bytes);
    const int nper = bytes / (P * 16);
    while (!mbar_try_wait(bar0 + 8 * stage, parity)) {
    }
    const uint8_t* sbuf = wbuf + stage * kChunk;
    if (do_mask) {
      // mask word of vector v of the item: v / LPW (nk % 32 == 0 on this path when the mask is wanted); mw0 = word of the
      // chunk's first vector (whole periods per chunk, and 32 | P, so it is a whole number of words)
      uint32_t* mw0 = a.mask + (size_t)c * ((size_t)a.S * (g.nk >> 5)) + ((size_t)j * (g.chunk_bytes >> 4)) / LPW;
      ijk_chunk_mask<T, SLOTS, R::kBatch5>(sbuf, nper, lane, a.thr, a.thr_f, mw0, acc, cnt);
    } else if (is_ql) {
      ijk_chunk<T, SLOTS, R::kBatch5, true, false>(sbuf, nper, lane, a.thr, nullptr, acc, cnt);
    } else {
      ijk_chunk<T, SLOTS, R::kBatch5, false, false>(sbuf, nper, lane, a.thr, nullptr, acc, cnt);
    }
    __syncwarp();  // every lane is done with the stage before it is refilled
    if (lane == 0 && pio < my_items) issue(stage);
    if (++stage == kStages) {
      stage = 0;
      parity ^= 1;
    }
    j += kWarps;
  }
  combine(true);
}

// The float32 form of the cloud test: (double)x > thr for a float32 x is the same predicate as x >= t with
// t = the smallest float32 strictly above thr (exact, including zeros, denormals and infinities; NaN on either side
// compares false both ways). One FSETP per value instead of a conversion-dependent DSETP.
float float_threshold(double thr) {
  if (thr != thr || thr == (double)INFINITY) return NAN;     // nothing is greater than NaN / +inf
  const float f = (float)thr;                               // round to nearest (may be +-inf)
  if ((double)f > thr) return f;
  return nextafterf(f, INFINITY);
}

bool fast_path(int dtype, long long S) {
  const long long slab_bytes = S * (dtype == SPC_F32 ? 4 : 8);
  return slab_bytes % 16 == 0 && slab_bytes >= 1024;
}

// Ring shapes. Production picks by slab size and element type (round-1 sweeps on B200, profiles/README.md): with the
// evict_normal L2 policy 96-128 KB in flight per SM in single-stage per-warp rings is the sweet spot - deeper rings
// (192 KB) cost 6-8 % of bandwidth, fewer than 8 warps cannot keep up with the float32->float64 conversions.
//   slab >= 8 KB : float32 16 warps x 1 x 8 KB, float64 12 warps x 1 x 8 KB
//   slab  = 4 KB : pairs of slabs, 12 warps x 1 x 8 KB (slab_reduce_tma_pair_kernel)     other slab < 8 KB : 24 warps x 1 x 4 KB
using RingF32 = Ring<16, 8192, 1, false, 1, 1>;
using RingF64 = Ring<12, 8192, 1, false, 1, 1>;
using RingSmall = Ring<24, 4096, 1, false, 1, 1>;
constexpr int kPairWarps = 12;

template <typename R>
constexpr size_t tma_smem() {
  return (size_t)R::kWarps * R::kStages * R::kChunk + (size_t)R::kWarps * R::kStages * 8;
}
template <typename T, typename R>
int configure_tma() {
  SPC_CUDA(cudaFuncSetAttribute(slab_reduce_tma_kernel<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem<R>()));
  return SPC_OK;
}
template <typename T, typename R>
int launch_tma(spc_handle h, const K1Args& a, cudaStream_t st) {
  const long long want = (a.total + R::kWarps - 1) / R::kWarps;
  const int grid = (int)std::min<long long>(h->num_sms, std::max<long long>(want, 1));
  slab_reduce_tma_kernel<T, R><<<grid, R::kWarps * 32, tma_smem<R>(), st>>>(a);
  return SPC_OK;
}

template <int W, int ST>
constexpr size_t pair_smem() {
  return (size_t)W * ST * 2 * kSubBytes + (size_t)W * ST * 8;
}
template <typename T, int W, int ST = 1>
int configure_tma_pair() {
  SPC_CUDA(cudaFuncSetAttribute(slab_reduce_tma_pair_kernel<T, W, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)pair_smem<W, ST>()));
  return SPC_OK;
}
template <typename T, int W, int ST = 1>
int launch_tma_pair(spc_handle h, const K1Args& a, cudaStream_t st) {
  const long long want = ((a.total >> 1) + W - 1) / W;
  const int grid = (int)std::min<long long>(h->num_sms, std::max<long long>(want, 1));
  slab_reduce_tma_pair_kernel<T, W, ST><<<grid, W * 32, pair_smem<W, ST>(), st>>>(a);
  return SPC_OK;
}

#ifdef SPC_TUNING
// Sweep variants of libspcpl_b200_tune.so (tools/k1_probe.py, spc_tune_k1); they document the round-1 sweeps and are
// not compiled into the production library. X(id, warps, chunk, stages, blocked); Y adds (L2 hint, sub-blocks per trip):
// hint 0 = evict_first, 1 = evict_normal (production), 2 = no cache hint, 3 = evict_last.
#define SPC_K1_VARIANTS(X)                                                               \
  X(1, 12, 8192, 1, false) X(2, 24, 4096, 1, false) X(3, 16, 4096, 3, false) X(4, 8, 8192, 2, false) \
  X(5, 8, 16384, 1, false) X(6, 16, 4096, 2, false) X(7, 12, 8192, 2, false) X(8, 12, 8192, 1, true)
#define SPC_K1_VARIANTS2(Y) Y(10, 12, 8192, 1, false, 0, 1) Y(11, 12, 8192, 1, false, 1, 2) Y(12, 12, 8192, 1, false, 2, 1) \
  Y(14, 12, 8192, 1, false, 3, 1) Y(15, 24, 4096, 1, false, 0, 1) \
  Y(16, 8, 8192, 3, false, 1, 1) Y(17, 8, 8192, 2, false, 1, 2) Y(18, 8, 8192, 2, false, 2, 1) Y(19, 16, 8192, 1, false, 1, 1) \
  Y(20, 8, 8192, 2, false, 0, 1) Y(21, 6, 16384, 2, false, 1, 1)
#endif

// chunk bytes of the ring that will serve this slab size / element type
int k1_chunk_bytes(spc_handle h, int slab_bytes, int esize) {
#ifdef SPC_TUNING
  switch (h->k1_variant) {
#define X(id, w, c, s, b) case id: return c;
    SPC_K1_VARIANTS(X)
#undef X
#define Y(id, w, c, s, b, hint, un) case id: return c;
    SPC_K1_VARIANTS2(Y)
#undef Y
    default: break;
  }
#endif
  (void)h; (void)esize;
  return slab_bytes < 8192 ? 4096 : 8192;
}

template <typename T>
int launch_kji(spc_handle h, const K1Args& a, bool fast, cudaStream_t st) {
  int rc = SPC_OK;
  if (!fast) {
    const long long want = (a.total + 7) / 8;
    const int grid = (int)std::min<long long>((long long)h->num_sms * 8, std::max<long long>(want, 1));
    slab_reduce_generic_kernel<T><<<grid, 256, 0, st>>>(a);
  } else
#ifdef SPC_TUNING
  if (h->k1_variant != 0 && !(a.slab_bytes == kSubBytes && a.per_field % 2 == 0 && (h->k1_variant == 9 || (h->k1_variant >= 22 && h->k1_variant <= 25)))) {
    switch (h->k1_variant) {
#define X(id, w, c, s, b) case id: rc = launch_tma<T, Ring<w, c, s, b>>(h, a, st); break;
      SPC_K1_VARIANTS(X)
#undef X
#define Y(id, w, c, s, b, hint, un) case id: rc = launch_tma<T, Ring<w, c, s, b, hint, un>>(h, a, st); break;
      SPC_K1_VARIANTS2(Y)
#undef Y
      default: rc = SPC_ERR_ARG; spc::set_error("unknown K1 variant %d", h->k1_variant); break;
    }
  } else if (h->k1_variant == 22) {
    rc = launch_tma_pair<T, 16>(h, a, st);
  } else if (h->k1_variant == 23) {
    rc = launch_tma_pair<T, 8, 2>(h, a, st);
  } else if (h->k1_variant == 24) {
    rc = launch_tma_pair<T, 12, 2>(h, a, st);
  } else if (h->k1_variant == 25) {
    rc = launch_tma_pair<T, 6, 3>(h, a, st);
  } else
#endif
  if (a.slab_bytes == kSubBytes && a.per_field % 2 == 0) {
    rc = launch_tma_pair<T, kPairWarps>(h, a, st);          // 4 KB slabs in pairs
  } else if (a.slab_bytes < 8192) {
    rc = launch_tma<T, RingSmall>(h, a, st);
  } else {
    rc = launch_tma<T, std::conditional_t<sizeof(T) == 4, RingF32, RingF64>>(h, a, st);
  }
  if (rc) return rc;
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

// production IJK ring (round-2 sweep, profiles/README.md): 8 warps x 2 stages x 8 KB keeps a copy in flight while a warp
// reduces the other stage - the heavier ql chunks (flags, counts, mask) no longer stall the stream: 6.57 -> 6.95 TB/s with
// the mask, 6.90 -> 7.19 TB/s without (the 12-warp single-stage ring of round 1 is tuning variant 7)
using IjkProd = IjkRing<8, 8192, 3, 2>;

template <typename T, int SLOTS, typename R>
constexpr size_t ijk_smem() {
  constexpr int VEC = 16 / (int)sizeof(T), P = SLOTS * 32;
  return (size_t)R::kWarps * R::kStages * R::kChunk + (size_t)R::kWarps * P * VEC * (sizeof(double) + sizeof(int)) +
         (size_t)R::kWarps * R::kStages * 8 + 16;
}
template <typename T, int SLOTS, typename R>
int configure_ijk_tma() {
  SPC_CUDA(cudaFuncSetAttribute(slab_reduce_ijk_tma_kernel<T, SLOTS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)ijk_smem<T, SLOTS, R>()));
  return SPC_OK;
}
template <typename T, int SLOTS, typename R>
int launch_ijk_tma(spc_handle h, const K1Args& a, IjkArgs g, cudaStream_t st) {
  constexpr int P = SLOTS * 32;
  static_assert(P * 16 <= R::kChunk, "a period must fit one ring stage");
  g.chunk_bytes = (R::kChunk / (P * 16)) * (P * 16);
  g.nch = (int)((g.item_bytes + g.chunk_bytes - 1) / g.chunk_bytes);
  const int grid = (int)std::min<long long>(h->num_sms, g.nitems);
  slab_reduce_ijk_tma_kernel<T, SLOTS, R><<<grid, R::kWarps * 32, ijk_smem<T, SLOTS, R>(), st>>>(a, g);
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

template <typename T, int SLOTS>
int launch_ijk_tma_variant(spc_handle h, const K1Args& a, const IjkArgs& g, cudaStream_t st) {
#ifdef SPC_TUNING
  switch (h->ijk_variant) {       // (warps, chunk, periods per pass, stages)
    case 2: return launch_ijk_tma<T, SLOTS, IjkRing<12, 8192, 1>>(h, a, g, st);
    case 3: return launch_ijk_tma<T, SLOTS, IjkRing<8, 8192, 2, 2>>(h, a, g, st);
    case 4: return launch_ijk_tma<T, SLOTS, IjkRing<6, 8192, 2, 3>>(h, a, g, st);
    case 6: return launch_ijk_tma<T, SLOTS, IjkRing<10, 8192, 2, 1>>(h, a, g, st);
    case 7: return launch_ijk_tma<T, SLOTS, IjkRing<12, 8192, 2, 1>>(h, a, g, st);
    default: break;
  }
#endif
  return launch_ijk_tma<T, SLOTS, IjkProd>(h, a, g, st);
}

// Eligibility of the IJK TMA path: whole 16-byte vectors per point, a period of at most 5 x 32 vectors,
// whole periods per item, 16-byte aligned volumes, and (for the per-point mask) nk a multiple of 32.
template <typename T>
bool ijk_tma_plan(spc_handle h, const K1Args& a, int ncol, int nk, IjkArgs& g, int& slots) {
  constexpr int V = 16 / (int)sizeof(T);
  if (h->ijk_variant == 1 || nk % V != 0) return false;   // ijk_variant 1 (tuning build only): always the CTA-per-item kernel
  const int nkv = nk / V;
  int gcd = nkv, b = 32;
  while (b) { const int t = gcd % b; gcd = b; b = t; }
  slots = nkv / gcd;
  const int rows_per_period = 32 / gcd;
  if (slots > 5 || a.S % rows_per_period != 0) return false;
  if (a.mask && nk % 32 != 0) return false;
  const void* vols[SPC_NFIELDS] = {a.v0, a.v1, a.v2, a.v3, a.v4};
  for (int f = 0; f < SPC_NFIELDS; ++f)
    if (reinterpret_cast<uintptr_t>(vols[f]) % 16 != 0) return false;
  g.nk = nk; g.ncol = ncol; g.nkv = nkv;
  g.item_bytes = (long long)a.S * nk * (long long)sizeof(T);
  g.nitems = (long long)SPC_NFIELDS * ncol;
  g.want_mask = a.mask != nullptr;
  return true;
}

template <typename T>
int launch_ijk(spc_handle h, const K1Args& a, int ncol, int nk, cudaStream_t st) {
  constexpr int V = 16 / sizeof(T);
  IjkArgs g;
  int slots = 0;
  if (ijk_tma_plan<T>(h, a, ncol, nk, g, slots)) {
    switch (slots) {
      case 1: return launch_ijk_tma_variant<T, 1>(h, a, g, st);
      case 2: return launch_ijk_tma_variant<T, 2>(h, a, g, st);
      case 3: return launch_ijk_tma_variant<T, 3>(h, a, g, st);
      case 4: return launch_ijk_tma_variant<T, 4>(h, a, g, st);
      default: return launch_ijk_tma_variant<T, 5>(h, a, g, st);
    }
  }
  bool vec = (nk % V == 0);
  const void* vols[SPC_NFIELDS] = {a.v0, a.v1, a.v2, a.v3, a.v4};
  for (int f = 0; f < SPC_NFIELDS && vec; ++f) vec = (reinterpret_cast<uintptr_t>(vols[f]) % 16 == 0);
  const int nkv = vec ? nk / V : nk;
  SPC_REQUIRE(nkv <= 1024, SPC_ERR_UNSUPPORTED, "spc_slab_reduce: nk=%d too large for the IJK layout kernel", nk);
  int R = std::max(1, std::min(a.S, 512 / nkv));
  const int kw = (nk + 31) / 32;
  auto smem_for = [&](int r) { return (size_t)r * nk * (sizeof(double) + sizeof(int)) + (size_t)r * kMaskRows * kw * sizeof(uint32_t); };
  size_t smem = smem_for(R);
  while (smem > 48 * 1024 && R > 1) {
    --R;
    smem = smem_for(R);
  }
  SPC_REQUIRE(smem <= 48 * 1024, SPC_ERR_UNSUPPORTED, "spc_slab_reduce: nk=%d too large for the IJK layout kernel", nk);
  const int threads = ((nkv * R + 31) / 32) * 32;
  const int grid = SPC_NFIELDS * ncol;
  if (vec)
    slab_reduce_ijk_kernel<T, V><<<grid, threads, smem, st>>>(a, nk, ncol, nkv, R);
  else
    slab_reduce_ijk_kernel<T, 1><<<grid, threads, smem, st>>>(a, nk, ncol, nkv, R);
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

template <typename T>
int configure_all() {
  int rc = SPC_OK;
  if ((rc = configure_tma<T, RingF32>())) return rc;
  if ((rc = configure_tma<T, RingF64>())) return rc;
  if ((rc = configure_tma<T, RingSmall>())) return rc;
  if ((rc = configure_tma_pair<T, kPairWarps>())) return rc;
  if ((rc = configure_ijk_tma<T, 1, IjkProd>())) return rc;
  if ((rc = configure_ijk_tma<T, 2, IjkProd>())) return rc;
  if ((rc = configure_ijk_tma<T, 3, IjkProd>())) return rc;
  if ((rc = configure_ijk_tma<T, 4, IjkProd>())) return rc;
  if ((rc = configure_ijk_tma<T, 5, IjkProd>())) return rc;
#ifdef SPC_TUNING
#define X(id, w, c, s, b) if ((rc = configure_tma<T, Ring<w, c, s, b>>())) return rc;
  SPC_K1_VARIANTS(X)
#undef X
#define Y(id, w, c, s, b, hint, un) if ((rc = configure_tma<T, Ring<w, c, s, b, hint, un>>())) return rc;
  SPC_K1_VARIANTS2(Y)
#undef Y
  if ((rc = configure_tma_pair<T, 16>())) return rc;
  if ((rc = configure_tma_pair<T, 8, 2>())) return rc;
  if ((rc = configure_tma_pair<T, 12, 2>())) return rc;
  if ((rc = configure_tma_pair<T, 6, 3>())) return rc;
#define SPC_IJK_CFG(...)                                                   \
  if ((rc = configure_ijk_tma<T, 1, IjkRing<__VA_ARGS__>>())) return rc;   \
  if ((rc = configure_ijk_tma<T, 2, IjkRing<__VA_ARGS__>>())) return rc;   \
  if ((rc = configure_ijk_tma<T, 3, IjkRing<__VA_ARGS__>>())) return rc;   \
  if ((rc = configure_ijk_tma<T, 4, IjkRing<__VA_ARGS__>>())) return rc;   \
  if ((rc = configure_ijk_tma<T, 5, IjkRing<__VA_ARGS__>>())) return rc;
  SPC_IJK_CFG(12, 8192, 1)
  SPC_IJK_CFG(8, 8192, 2, 2)
  SPC_IJK_CFG(6, 8192, 2, 3)
  SPC_IJK_CFG(10, 8192, 2, 1)
  SPC_IJK_CFG(12, 8192, 2, 1)
#undef SPC_IJK_CFG
#endif
  return rc;
}

}  // namespace

// Called once per handle (spc_create, on the handle's device): every streaming kernel of this file is opted in to its
// dynamic shared memory there, so the launch paths carry no lazily initialised state and are re-entrant.
int spc::k1_configure(spc_ctx*) {
  const int rc = configure_all<float>();
  return rc ? rc : configure_all<double>();
}

extern "C" {

#ifdef SPC_TUNING
// Tuning hook of libspcpl_b200_tune.so (tools/k1_probe.py), not part of the ABI: selects the TMA ring shape of this handle.
int spc_tune_k1(spc_handle h, int variant) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  if (variant >= 100) h->ijk_variant = variant - 100;
  else h->k1_variant = variant;
  return SPC_OK;
}
#endif

size_t spc_mask_words_per_column(int dtype, int layout, int nx, int ny, int nk) {
  if (nx <= 0 || ny <= 0 || nk <= 0 || (layout != SPC_LAYOUT_KJI && layout != SPC_LAYOUT_IJK)) return 0;
  const long long S = (long long)nx * ny;
  if (layout == SPC_LAYOUT_IJK) return (size_t)S * ((nk + 31) / 32);  // [S][ceil(nk/32)] words per column
  if (fast_path(dtype, S)) {
    const long long slab_bytes = S * (dtype == SPC_F32 ? 4 : 8);
    return (size_t)((slab_bytes + kSubBytes - 1) / kSubBytes) * 32 * nk;
  }
  return (size_t)((S + 31) / 32) * nk;
}

int spc_slab_reduce(spc_handle h, const void* const vol[5], int dtype, int layout, int ncol, int nx, int ny, int nk,
                    double ql_thresh, double* prof, int32_t* cnt, uint32_t* mask, void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(dtype == SPC_F32 || dtype == SPC_F64, SPC_ERR_ARG, "spc_slab_reduce: bad dtype %d", dtype);
  SPC_REQUIRE(layout == SPC_LAYOUT_KJI || layout == SPC_LAYOUT_IJK, SPC_ERR_ARG, "spc_slab_reduce: bad layout %d", layout);
  SPC_REQUIRE(ncol >= 0 && nx > 0 && ny > 0 && nk > 0, SPC_ERR_ARG, "spc_slab_reduce: bad shape ncol=%d nx=%d ny=%d nk=%d",
              ncol, nx, ny, nk);
  const long long S = (long long)nx * ny;
  SPC_REQUIRE(S * (dtype == SPC_F32 ? 4 : 8) < (1ll << 31), SPC_ERR_UNSUPPORTED, "spc_slab_reduce: slab too large");
  if (ncol == 0) return SPC_OK;
  SPC_REQUIRE(vol && prof, SPC_ERR_ARG, "spc_slab_reduce: vol/prof is NULL");
  for (int f = 0; f < SPC_NFIELDS; ++f) SPC_REQUIRE(vol[f] != nullptr, SPC_ERR_ARG, "spc_slab_reduce: vol[%d] is NULL", f);
  spc::DeviceGuard guard(h->device);
  K1Args a;
  a.v0 = vol[0]; a.v1 = vol[1]; a.v2 = vol[2]; a.v3 = vol[3]; a.v4 = vol[4];
  a.prof = prof;
  a.cnt = cnt;
  a.mask = mask;
  a.thr = ql_thresh;
  a.thr_f = float_threshold(ql_thresh);
  a.per_field = (long long)ncol * nk;
  a.total = a.per_field * SPC_NFIELDS;
  a.S = (int)S;
  a.slab_bytes = (int)(S * (dtype == SPC_F32 ? 4 : 8));
  const int chunk = k1_chunk_bytes(h, a.slab_bytes, dtype == SPC_F32 ? 4 : 8);
  a.nch = (a.slab_bytes + chunk - 1) / chunk;
  a.nsub = (a.slab_bytes + kSubBytes - 1) / kSubBytes;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == SPC_LAYOUT_IJK) {
    return dtype == SPC_F32 ? launch_ijk<float>(h, a, ncol, nk, st) : launch_ijk<double>(h, a, ncol, nk, st);
  }
  const bool fast = fast_path(dtype, S);
  if (fast) {
    for (int f = 0; f < SPC_NFIELDS; ++f)
      SPC_REQUIRE(reinterpret_cast<uintptr_t>(vol[f]) % 16 == 0, SPC_ERR_ALIGN,
                  "spc_slab_reduce: vol[%d] must be 16-byte aligned for the TMA bulk path", f);
  }
  return dtype == SPC_F32 ? launch_kji<float>(h, a, fast, st) : launch_kji<double>(h, a, fast, st);
}

}  // extern "C"
