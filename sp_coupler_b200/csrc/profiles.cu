// K2 gcm_to_les, K3 les_to_gcm and the sputils batch helpers (interp / searchsorted / exner).
//
// One thread block per superparameterized column; the column's GCM profiles are converted once
// into shared memory (heights, theta_l, q_t, reversed to ascending height) and every LES / GCM
// level is then handled by one thread: bracket search in shared memory + all interpolations +
// the relaxation forcing, fused. All arithmetic is float64 and this file is compiled with
// --fmad=false so that float64 results are bit-identical to numpy's wherever numpy's own
// operations are correctly rounded (everything except pow, which CUDA evaluates to <= 2 ulp).
#include "spc_common.cuh"

namespace {

using namespace spc;

constexpr int kThreads = 256;
// (spc_ctx.k2_threads / k3_threads / proj_threads: tuning overrides, always 0 in the production library)

// Threads per column CTA: one per level of the longer of the two profiles, whole warps, at most kThreads.
// (256 threads for 137 / 160 levels left three of eight warps idle and capped the resident useful warps.)
int column_threads(int tuned, int n) {
  if (tuned) return tuned;
  return std::min(kThreads, std::max(32, (n + 31) & ~31));
}
constexpr double kC = rv / rd - 1;          // spcpl.py:175
constexpr double kExp = rd / cp;            // sputils.py:29
constexpr double kIExp = -rd / cp;          // sputils.py:34

struct GcmPtrs {
  const void *U, *V, *T, *SH, *QL, *QI, *Pfull, *A, *Zgfull, *Phalf, *Zghalf;
  const void *Z0M, *Z0H, *QLflux, *QIflux, *SHflux, *TLflux, *TSflux;
  int ncol, nlev;
};

GcmPtrs to_ptrs(const spc_gcm_cols* g) {
  GcmPtrs p;
  p.U = g->U; p.V = g->V; p.T = g->T; p.SH = g->SH; p.QL = g->QL; p.QI = g->QI; p.Pfull = g->Pfull; p.A = g->A;
  p.Zgfull = g->Zgfull; p.Phalf = g->Phalf; p.Zghalf = g->Zghalf;
  p.Z0M = g->Z0M; p.Z0H = g->Z0H; p.QLflux = g->QLflux; p.QIflux = g->QIflux; p.SHflux = g->SHflux;
  p.TLflux = g->TLflux; p.TSflux = g->TSflux;
  p.ncol = g->ncol; p.nlev = g->nlev;
  return p;
}

template <typename T>
__device__ __forceinline__ double ld(const void* p, size_t i) {
  return (double)__ldg(static_cast<const T*>(p) + i);
}
template <typename T>
__device__ __forceinline__ void st(void* p, size_t i, double v) {
  if (p) static_cast<T*>(p)[i] = (T)v;
}

// ------------------------------------------------------------------------------------------ K2
struct K2Args {
  GcmPtrs g;
  const double *zf, *zh, *les_prof;
  const void* ps_les;
  spc_les_forcing o;
  double dt, factor;
  int nk, couple_surface;
};

// PREFETCH: request the operands of the thread's first LES level before phase 1 (a shorter dependent chain, 14 more
// registers): pays when the launch is a single wave of CTAs, i.e. latency-bound; larger launches are bound by the number of
// resident CTAs and run the leaner variant.
template <typename T, bool PREFETCH>
__global__ void __launch_bounds__(kThreads) gcm_to_les_kernel(const K2Args a) {
  extern __shared__ __align__(16) double sm[];
  const int nlev = a.g.nlev, nk = a.nk, ncol = a.g.ncol;
  const int c = blockIdx.x;
  double* Zf = sm;               // ascending (reversed GCM order), spcpl.py:224-228
  double* thl = Zf + nlev;
  double* qt = thl + nlev;
  double* ql = qt + nlev;
  double* u = ql + nlev;
  double* v = u + nlev;
  double* zh_s = v + nlev;       // [nk] LES half levels (only when the cloud-slab mapping is wanted)
  const size_t b = (size_t)c * nlev, bh = (size_t)c * (nlev + 1);
  const double zs = ld<T>(a.g.Zghalf, bh + nlev);   // Zghalf[-1]
  const spc_les_forcing& o = a.o;
  const bool want_idx = o.slab_idx && a.zh;         // block-uniform
  // operands of this thread's first LES level (phase 2) are requested now, so that they are in flight during phase 1
  const size_t pf = (size_t)ncol * nk;   // field stride of les_prof [5][ncol][nk]
  const int k_first = threadIdx.x;
  double x_first = 0.0, lp_first[SPC_NFIELDS] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (PREFETCH && k_first < nk) {
    x_first = __ldg(a.zf + k_first);
    if (a.les_prof) {
#pragma unroll
      for (int f = 0; f < SPC_NFIELDS; ++f) lp_first[f] = __ldg(a.les_prof + f * pf + (size_t)c * nk + k_first);
    }
  }
  if (want_idx) {
    for (int k = threadIdx.x; k < nk; k += blockDim.x) zh_s[k] = __ldg(a.zh + k);
    __syncthreads();
  }

  for (int l = threadIdx.x; l <= nlev; l += blockDim.x) {
    const double Zh = (ld<T>(a.g.Zghalf, bh + l) - zs) / grav;          // spcpl.py:197
    st<T>(o.Zh, bh + l, Zh);
    if (l < nlev && want_idx) {
      // searchsorted(zh, Zh, side="right")[:-1][::-1]  (spcpl.py:26,764)
      o.slab_idx[b + (nlev - 1 - l)] = upper_bound(zh_s, nk, Zh);
    }
    if (l == nlev) break;
    const double Tl = ld<T>(a.g.T, b + l), SH = ld<T>(a.g.SH, b + l);
    const double QL = ld<T>(a.g.QL, b + l), QI = ld<T>(a.g.QI, b + l);
    const double Pf = ld<T>(a.g.Pfull, b + l);
    const double Tv = Tl * (1 + kC * SH - (QL + QI));                    // spcpl.py:176
    const double zf_l = (ld<T>(a.g.Zgfull, b + l) - zs) / grav;         // spcpl.py:198
    const double thl_ = (Tl - (rlv * (QL + QI)) / cp) * pow(Pf / pref0, kIExp);  // spcpl.py:214
    const double qt_ = SH + QL + QI;                                     // spcpl.py:215
    const int r = nlev - 1 - l;
    Zf[r] = zf_l;
    thl[r] = thl_;
    qt[r] = qt_;
    ql[r] = QL;
    u[r] = ld<T>(a.g.U, b + l);
    v[r] = ld<T>(a.g.V, b + l);
    st<T>(o.Tv, b + l, Tv);
    st<T>(o.Zf, b + l, zf_l);
    st<T>(o.THL, b + l, thl_);
    st<T>(o.QT, b + l, qt_);
  }
  if (threadIdx.x == 0) {
    const double ps = ld<T>(a.g.Phalf, bh + nlev);                       // Ph[-1], spcpl.py:246
    st<T>(o.ps, c, ps);
    if (o.f_ps) st<T>(o.f_ps, c, a.factor * (ps - ld<T>(a.ps_les, c)) / a.dt);   // spcpl.py:332
    if (a.couple_surface) {                                              // spcpl.py:136-167
      const double rho = ps / (rd * ld<T>(a.g.T, b + nlev - 1));
      const double wqt = -(ld<T>(a.g.QLflux, c) + ld<T>(a.g.QIflux, c) + ld<T>(a.g.SHflux, c)) / rho;
      const double wthl = -ld<T>(a.g.TSflux, c) * pow(ps / pref0, kIExp) / (cp * rho);
      st<T>(o.z0m, c, ld<T>(a.g.Z0M, c));
      st<T>(o.z0h, c, ld<T>(a.g.Z0H, c));
      st<T>(o.wthl, c, wthl);
      st<T>(o.wqt, c, wqt);
    }
  }
  __syncthreads();

  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    const bool first = PREFETCH && (k == k_first);
    const double x = first ? x_first : __ldg(a.zf + k);
    const int j = upper_bound(Zf, nlev, x) - 1;
    const size_t i = (size_t)c * nk + k;
    const double thl_g = interp_at(Zf, thl, nlev, x, j);
    const double qt_g = interp_at(Zf, qt, nlev, x, j);
    const double ql_g = interp_at(Zf, ql, nlev, x, j);
    const double u_g = interp_at(Zf, u, nlev, x, j);
    const double v_g = interp_at(Zf, v, nlev, x, j);
    if (o.bracket) o.bracket[i] = j;
    st<T>(o.thl, i, thl_g);
    st<T>(o.qt, i, qt_g);
    st<T>(o.u, i, u_g);
    st<T>(o.v, i, v_g);
    st<T>(o.ql_ref, i, ql_g);
    if (a.les_prof) {                                                    // spcpl.py:328-333
      double lp[SPC_NFIELDS];
#pragma unroll
      for (int f = 0; f < SPC_NFIELDS; ++f) lp[f] = first ? lp_first[f] : __ldg(a.les_prof + f * pf + i);
      st<T>(o.f_u, i, a.factor * (u_g - lp[SPC_U]) / a.dt);
      st<T>(o.f_v, i, a.factor * (v_g - lp[SPC_V]) / a.dt);
      st<T>(o.f_thl, i, a.factor * (thl_g - lp[SPC_THL]) / a.dt);
      st<T>(o.f_qt, i, a.factor * (qt_g - lp[SPC_QT]) / a.dt);
      st<T>(o.f_ql, i, a.factor * (ql_g - lp[SPC_QL]) / a.dt);
    }
  }
}

// ------------------------------------------------------------------------------------------ K3
struct K3Args {
  GcmPtrs g;
  const double *zf, *zh;
  spc_les_prof les;
  spc_gcm_tend o;
  double dt, factor;
  int nk, conservative;
  int project_mw;       // > 0: the KJI cloud projection runs in this kernel's prologue; mask words per level
  // fused gather / host exchange targets (spc_gcm_tend): n_peers targets x n_bufs buffer sets
  void* peers[2 * SPC_MAX_PEERS];
  int n_peers, n_bufs;
  size_t peer_off;      // element offset of this rank's block inside a target buffer
  // completion protocol (spc_gcm_tend.sync)
  uint32_t* sync;
  uint32_t* signal[SPC_MAX_PEERS + 1];
  int n_signal, sync_slot, n_wait;
};

// sputils.integral, weighted branch (sputils.py:94-161), a <= b guaranteed by the caller
// (a = Zh[i+1] < b = Zh[i]); z has n entries, q/w are cell values on [z[i], z[i+1]].
__device__ double integral_w(double a, double b, const double* z, int n, const double* q, const double* w) {
  int ia = 0;
  while (ia + 1 < n - 1 && z[ia + 1] < a) ++ia;                         // sputils.py:123-124
  int ib = ia;
  while (ib + 1 < n - 1 && z[ib + 1] < b) ++ib;                         // sputils.py:126-127
  double S = 0.0, Sw = 0.0;
  for (int i = ia; i <= ib; ++i) {
    const double dz = z[i + 1] - z[i];
    S += w[i] * q[i] * dz;                                              // sputils.py:152
    Sw += w[i] * dz;                                                    // sputils.py:157
  }
  const double Sa = w[ia] * q[ia] * (a - z[ia]);                        // sputils.py:154
  const double Sb = w[ib] * q[ib] * (z[ib + 1] - b);                    // sputils.py:155
  const double Swa = w[ia] * (a - z[ia]);                               // sputils.py:159
  const double Swb = w[ib] * (z[ib + 1] - b);                           // sputils.py:160
  return (S - Sa - Sb) / (Sw - Swa - Swb);                              // sputils.py:161
}

// The same integral for NF profiles that share the edges z and the weights w (the seven interp_c calls of
// spcpl.py:480-486 differ only in q): the two cell searches, the weight sums and the edge lengths are
// evaluated once, and every field sees exactly the operations, in the order, of integral_w - the results are
// bit-identical to NF separate calls. q[f] are [n-1] cell values; z, w, q live in shared memory.
template <int NF>
__device__ __forceinline__ void integral_w_multi(double a, double b, const double* z, int n, const double* const (&q)[NF],
                                                 const double* w, double (&out)[NF]) {
  int ia = 0;
  while (ia + 1 < n - 1 && z[ia + 1] < a) ++ia;
  int ib = ia;
  while (ib + 1 < n - 1 && z[ib + 1] < b) ++ib;
  double S[NF], Sw = 0.0;
#pragma unroll
  for (int f = 0; f < NF; ++f) S[f] = 0.0;
  for (int i = ia; i <= ib; ++i) {
    const double dz = z[i + 1] - z[i], wi = w[i];
#pragma unroll
    for (int f = 0; f < NF; ++f) S[f] += wi * q[f][i] * dz;
    Sw += wi * dz;
  }
  const double da = a - z[ia], db = z[ib + 1] - b, wa = w[ia], wb = w[ib];
  const double den = Sw - wa * da - wb * db;
#pragma unroll
  for (int f = 0; f < NF; ++f) out[f] = (S[f] - wa * q[f][ia] * da - wb * q[f][ib] * db) / den;
}

// ---- projected cloud cover per GCM slab (les.get_cloudfraction(indices), spcpl.py:28,765) ----------
// Slab r of a column covers the LES levels [k0, k1), k1 = min(max(idx[0..r]), nk), k0 likewise for r-1
// (idx = searchsorted(zh, Zh, 'right')[:-1][::-1], spcpl.py:26,764, made monotone and clipped).
// KJI mask: one block per column. The mask layout is opaque but identical for every level, so a slab's
// projected count is popcount(OR over its levels), word by word. Slabs without a cloudy level (most of them:
// the K1 counts say which levels have cloud) get their zero from one thread each; the others are queued and
// taken one per WARP, which ORs the slab's cloudy levels with 128-bit loads (lane <-> word group, up to eight
// independent loads in flight), popcounts and writes one exact integer. No atomics on data, no dependent loads.
// Prologue shared by both mask layouts: kend[r] = min(running max of slab_idx, nk), live_k[k] = level k has a cloudy
// cell (K1 count != 0; all levels when no counts are given); slabs that are empty or cloud-free get their zero here,
// the others are queued. Returns the queue length (after a block barrier).
template <typename T>
__device__ __forceinline__ int cloud_slab_queue(const int32_t* slab_idx, const int32_t* cnt, int nk, int nlev, int c,
                                                int* kend, int* live_k, int* queue, int32_t* cntslab, T* A,
                                                int* cs_sh = nullptr) {
  __shared__ int wmax[32];
  __shared__ int nqueue;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int32_t* cn = cnt ? cnt + (size_t)c * nk : nullptr;
  if (threadIdx.x == 0) nqueue = 0;
  for (int k = threadIdx.x; k < nk; k += blockDim.x) live_k[k] = cn ? (__ldg(cn + k) != 0) : 1;
  // running maximum of slab_idx: warp-shuffle scan per 32 slabs, warp maxima through shared memory, carry per pass
  int carry = 0;
  for (int r0 = 0; r0 < nlev; r0 += blockDim.x) {      // block-uniform trip count
    const int r = r0 + threadIdx.x;
    int v = r < nlev ? __ldg(slab_idx + (size_t)c * nlev + r) : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v = max(v, up);
    }
    if (lane == 31) wmax[warp] = v;
    __syncthreads();
    int before = carry;
    for (int w = 0; w < warp; ++w) before = max(before, wmax[w]);
    if (r < nlev) kend[r] = min(max(v, before), nk);
    for (int w = 0; w < nwarps; ++w) carry = max(carry, wmax[w]);
    __syncthreads();
  }
  // one thread per slab: empty or cloud-free slabs are answered here, the rest are queued
  for (int r = threadIdx.x; r < nlev; r += blockDim.x) {
    const int k0 = r ? kend[r - 1] : 0, k1 = kend[r];
    bool any = false;
    for (int k = k0; k < k1; ++k) any = any || live_k[k];
    if (any) {
      queue[atomicAdd(&nqueue, 1)] = r;     // order is irrelevant: slabs are independent
    } else {
      if (cntslab) cntslab[(size_t)c * nlev + r] = 0;
      if (A) A[(size_t)c * nlev + r] = (T)0;
      if (cs_sh) cs_sh[r] = 0;
    }
  }
  __syncthreads();
  return nqueue;
}

// The queued (cloudy) slabs of one column, one per warp; results to global memory (cntslab / A, ascending slab
// order) and/or to the block's shared array cs_sh (K3's fused prologue). No barrier inside.
template <typename T>
__device__ __forceinline__ void project_kji_queued(const uint32_t* m, bool vec, int mw, int nq, const int* queue,
                                                   const int* kend, const int* live_k, int nlev, int c, double npts,
                                                   int32_t* cntslab, T* A, int* cs_sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int qi = warp; qi < nq; qi += nwarps) {
    const int r = queue[qi];
    const int k0 = r ? kend[r - 1] : 0, k1 = kend[r];
    int n = 0;
    if (vec) {
      constexpr int KB = 4, SEG = 2;
      const int nvr = mw >> 2;                                        // 16-byte vectors per level
      const bool packed = nvr < 32 && (32 % nvr) == 0;                // several levels per warp pass
      const int rpp = packed ? 32 / nvr : 1;
      const int roff = packed ? lane / nvr : 0, vpos = packed ? lane % nvr : lane;
      const uint4* mv = reinterpret_cast<const uint4*>(m);
      for (int v0 = 0; v0 < nvr; v0 += 32 * SEG) {
        uint4 acc[SEG];
#pragma unroll
        for (int sg = 0; sg < SEG; ++sg) acc[sg] = make_uint4(0u, 0u, 0u, 0u);
        for (int kb = k0; kb < k1; kb += KB * rpp) {
          uint4 v[KB][SEG];
#pragma unroll
          for (int j = 0; j < KB; ++j) {
            const int k = kb + j * rpp + roff;
            const bool live = k < k1 && live_k[k];
#pragma unroll
            for (int sg = 0; sg < SEG; ++sg) {
              const int p = v0 + sg * 32 + vpos;
              v[j][sg] = (live && p < nvr) ? __ldg(mv + (size_t)k * nvr + p) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
#pragma unroll
          for (int j = 0; j < KB; ++j)
#pragma unroll
            for (int sg = 0; sg < SEG; ++sg) {
              acc[sg].x |= v[j][sg].x;
              acc[sg].y |= v[j][sg].y;
              acc[sg].z |= v[j][sg].z;
              acc[sg].w |= v[j][sg].w;
            }
        }
        if (packed) {                                                 // OR the level rows that shared this pass
          for (int o = nvr; o < 32; o <<= 1) {
            acc[0].x |= __shfl_xor_sync(0xffffffffu, acc[0].x, o);
            acc[0].y |= __shfl_xor_sync(0xffffffffu, acc[0].y, o);
            acc[0].z |= __shfl_xor_sync(0xffffffffu, acc[0].z, o);
            acc[0].w |= __shfl_xor_sync(0xffffffffu, acc[0].w, o);
          }
          if (roff != 0) acc[0] = make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int sg = 0; sg < SEG; ++sg) n += __popc(acc[sg].x) + __popc(acc[sg].y) + __popc(acc[sg].z) + __popc(acc[sg].w);
      }
    } else {
      for (int w = lane; w < mw; w += 32) {
        uint32_t acc = 0u;
        for (int k = k0; k < k1; ++k)
          if (live_k[k]) acc |= __ldg(m + (size_t)k * mw + w);
        n += __popc(acc);
      }
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) {
      if (cntslab) cntslab[(size_t)c * nlev + r] = n;
      if (A) A[(size_t)c * nlev + r] = (T)((double)n / npts);
      if (cs_sh) cs_sh[r] = n;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cloud_project_kji_kernel(const uint32_t* mask, const int32_t* slab_idx,
                                                                const int32_t* cnt, int mw, int nk, int nlev, double npts,
                                                                int32_t* cntslab, T* A) {
  extern __shared__ __align__(16) double sm[];
  int* kend = reinterpret_cast<int*>(sm);  // [nlev] exclusive end level of every slab
  int* live_k = kend + nlev;               // [nk] 1 where the level has any cloudy cell
  int* queue = live_k + nk;                // [nlev] slabs that need the mask
  const int c = blockIdx.x;
  const int nq = cloud_slab_queue<T>(slab_idx, cnt, nk, nlev, c, kend, live_k, queue, cntslab, A);
  const bool vec = (mw & 3) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0;   // block-uniform
  project_kji_queued<T>(mask + (size_t)c * nk * mw, vec, mw, nq, queue, kend, live_k, nlev, c, npts, cntslab, A, nullptr);
}

// Same projection for the IJK-layout mask of K1: [S][kw] words per column, bit k%32 of word k/32
// of horizontal point `row` (the levels of a point are contiguous). One thread per point walks the
// slabs and tests its bit range; warp-reduced integer counts.
__device__ __forceinline__ void project_cloud_rows(const uint32_t* m, const int32_t* idx, int S, int kw, int nk, int nlev,
                                                   int* cslab) {
  const int S_pad = (S + 31) & ~31;
  for (int row = threadIdx.x; row < S_pad; row += blockDim.x) {
    const uint32_t* w = m + (size_t)row * kw;
    const bool valid = row < S;
    int k0 = 0;
    for (int r = 0; r < nlev && k0 < nk; ++r) {  // block-uniform trip count (idx is per column)
      const int k1 = min(max(__ldg(idx + r), k0), nk);
      bool any = false;
      if (valid && k1 > k0) {
        for (int q = k0 >> 5; q <= (k1 - 1) >> 5; ++q) {
          const int lo = max(k0 - (q << 5), 0), hi = min(k1 - (q << 5), 32);   // bit range [lo, hi) of word q
          const uint32_t range = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
          any = any || ((__ldg(w + q) & range) != 0u);
        }
      }
      const int n = __reduce_add_sync(0xffffffffu, any ? 1 : 0);
      if (n && (threadIdx.x & 31) == 0) atomicAdd(&cslab[r], n);
      k0 = k1;
    }
  }
}

// IJK mask, nk <= 256: one block per column. A thread keeps the mask words of ROWS horizontal points in registers
// (one coalesced read of the column's mask), then the block walks the queued (cloudy) slabs: the slab's bit range is
// tested against the 1-2 words it touches (the word loop is unrolled and skipped by a block-uniform test), the points
// with any cloud are counted with a warp reduction and one shared-memory add per warp and slab. Exact integers.
template <typename T, int KW, int ROWS>
__global__ void __launch_bounds__(256) cloud_project_ijk_reg_kernel(const uint32_t* mask, const int32_t* slab_idx,
                                                                    const int32_t* cnt, int kw, int S, size_t per_col, int nk,
                                                                    int nlev, double npts, int32_t* cntslab, T* A) {
  extern __shared__ __align__(16) double sm[];
  int* kend = reinterpret_cast<int*>(sm);
  int* live_k = kend + nlev;
  int* queue = live_k + nk;
  int* cq = queue + nlev;                  // [nlev] projected count of every queued slab
  const int c = blockIdx.x, lane = threadIdx.x & 31;
  const int nq = cloud_slab_queue<T>(slab_idx, cnt, nk, nlev, c, kend, live_k, queue, cntslab, A);
  for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) cq[qi] = 0;
  __syncthreads();
  const uint32_t* m = mask + (size_t)c * per_col;
  for (int row0 = 0; row0 < S && nq > 0; row0 += blockDim.x * ROWS) {      // block-uniform
    uint32_t w[ROWS][KW];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const int row = row0 + i * blockDim.x + threadIdx.x;
#pragma unroll
      for (int q = 0; q < KW; ++q) w[i][q] = (row < S && q < kw) ? __ldg(m + (size_t)row * kw + q) : 0u;
    }
    for (int qi = 0; qi < nq; ++qi) {
      const int r = queue[qi];
      const int k0 = r ? kend[r - 1] : 0, k1 = kend[r];
      const int q_lo = k0 >> 5, q_hi = (k1 - 1) >> 5;
      uint32_t any[ROWS];
#pragma unroll
      for (int i = 0; i < ROWS; ++i) any[i] = 0u;
#pragma unroll
      for (int q = 0; q < KW; ++q) {
        if (q >= q_lo && q <= q_hi) {                                       // block-uniform
          const int lo = max(k0 - (q << 5), 0), hi = min(k1 - (q << 5), 32);   // bit range [lo, hi) of word q
          const uint32_t range = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
#pragma unroll
          for (int i = 0; i < ROWS; ++i) any[i] |= w[i][q] & range;
        }
      }
      int n = 0;
#pragma unroll
      for (int i = 0; i < ROWS; ++i) n += (any[i] != 0u);
      n = __reduce_add_sync(0xffffffffu, n);
      if (lane == 0 && n) atomicAdd(&cq[qi], n);
    }
  }
  __syncthreads();
  for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) {
    const int r = queue[qi];
    if (cntslab) cntslab[(size_t)c * nlev + r] = cq[qi];
    if (A) A[(size_t)c * nlev + r] = (T)((double)cq[qi] / npts);
  }
}

// IJK mask, any nk: one block per column (project_cloud_rows above).
template <typename T>
__global__ void __launch_bounds__(kThreads) cloud_project_ijk_kernel(const uint32_t* mask, const int32_t* slab_idx, int kw, int S,
                                                                     size_t per_col, int nk, int nlev, double npts,
                                                                     int32_t* cntslab, T* A) {
  extern __shared__ __align__(16) double sm[];
  int* cslab = reinterpret_cast<int*>(sm);
  const int c = blockIdx.x;
  for (int l = threadIdx.x; l < nlev; l += kThreads) cslab[l] = 0;
  __syncthreads();
  project_cloud_rows(mask + (size_t)c * per_col, slab_idx + (size_t)c * nlev, S, kw, nk, nlev, cslab);
  __syncthreads();
  for (int l = threadIdx.x; l < nlev; l += kThreads) {
    if (cntslab) cntslab[(size_t)c * nlev + l] = cslab[l];
    if (A) A[(size_t)c * nlev + l] = (T)((double)cslab[l] / npts);
  }
}

// Launches the projection for either mask layout (shared by spc_cloud_fraction and spc_les_to_gcm).
template <typename T>
int launch_cloud_projection(spc_handle h, const uint32_t* mask, const int32_t* slab_idx, const int32_t* cnt, int vol_dtype,
                            int layout, int nx, int ny, int nk, int ncol, int nlev, int32_t* cntslab, T* A, cudaStream_t st) {
  const size_t per_col = spc_mask_words_per_column(vol_dtype, layout, nx, ny, nk);
  SPC_REQUIRE(per_col > 0, SPC_ERR_UNSUPPORTED, "no cloud mask format for this layout/shape");
  const double npts = (double)nx * (double)ny;
  if (layout == SPC_LAYOUT_KJI) {
    const size_t smem = ((size_t)2 * nlev + nk) * sizeof(int);
    SPC_REQUIRE(smem <= 48 * 1024, SPC_ERR_UNSUPPORTED, "nlev=%d, nk=%d too large", nlev, nk);
    cloud_project_kji_kernel<T><<<ncol, h->proj_threads ? h->proj_threads : (per_col / nk <= 128 ? 64 : 256), smem, st>>>(mask, slab_idx, cnt, (int)(per_col / nk), nk, nlev, npts, cntslab, A);
  } else {
    const int kw = (nk + 31) / 32, S = nx * ny;
    const size_t smem_reg = ((size_t)3 * nlev + nk) * sizeof(int);
    if (kw <= 8 && smem_reg <= 48 * 1024) {
#define SPC_IJK_PROJ(KW, ROWS)                                                                                          \
  cloud_project_ijk_reg_kernel<T, KW, ROWS><<<ncol, 256, smem_reg, st>>>(mask, slab_idx, cnt, kw, S, per_col, nk, nlev, npts, \
                                                                         cntslab, A)
      if (kw <= 2) SPC_IJK_PROJ(2, 16);
      else if (kw <= 4) SPC_IJK_PROJ(4, 16);
      else if (kw == 5) SPC_IJK_PROJ(5, 16);
      else SPC_IJK_PROJ(8, 8);
#undef SPC_IJK_PROJ
    } else {
      const size_t smem = (size_t)nlev * sizeof(int);
      SPC_REQUIRE(smem <= 48 * 1024, SPC_ERR_UNSUPPORTED, "nlev=%d too large", nlev);
      cloud_project_ijk_kernel<T><<<ncol, kThreads, smem, st>>>(mask, slab_idx, kw, S, per_col, nk, nlev, npts, cntslab, A);
    }
  }
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

// system-scope release store / acquire load of a completion flag (peer GPU memory over NVLink or pinned host memory)
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// K3. One CTA per column:
//   prologue  (KJI mask) the projected cloud count of every GCM slab, straight from K1's bit mask into shared memory
//   stage     LES profiles + reversed GCM heights / pressures in shared memory
//   levels    one thread per GCM level: bracket search, 7 interpolations (or the mass-weighted integrals), 7 tendencies,
//             stored into the packed block and into every remote target (peer GPUs over NVLink / pinned host memory)
//   epilogue  (sync) last CTA of the launch: signal flags, wait for the peers, advance the epoch
template <typename T>
__global__ void __launch_bounds__(kThreads, 4) les_to_gcm_kernel(const K3Args a) {
  extern __shared__ __align__(16) double sm[];
  const int nlev = a.g.nlev, nk = a.nk, ncol = a.g.ncol;
  const int c = blockIdx.x;
  double* ZfA = sm;              // [nlev] ascending heights  (Zf[::-1])
  double* PfA = ZfA + nlev;      // [nlev] Pf[::-1]
  double* zf = PfA + nlev;       // [nk]
  double* t_d = zf + nk;         // 7 LES profiles [nk] each (spcpl.py:471-477)
  double* qt_d = t_d + nk;
  double* ql_d = qt_d + nk;
  double* qlw_d = ql_d + nk;
  double* qli_d = qlw_d + nk;
  double* u_d = qli_d + nk;
  double* v_d = u_d + nk;
  double* rho = v_d + nk;        // [nk]   (conservative only)
  double* ZhD = rho + nk;        // [nlev+1] descending half-level heights (conservative only)
  double* zh_s = ZhD + nlev + 1; // [nk]   LES half levels (conservative only)
  double* gT = zh_s + nk;        // 7 GCM profiles [nlev] each, GCM order (the X of f_X = factor (x_les - X) / dt)
  double* gSH = gT + nlev;
  double* gQL = gSH + nlev;
  double* gQI = gQL + nlev;
  double* gU = gQI + nlev;
  double* gV = gU + nlev;
  double* gA = gV + nlev;
  int* cs_sh = reinterpret_cast<int*>(gA + nlev);   // [nlev] projected cloud counts, ascending slabs (fused projection)
  int* kend = cs_sh + nlev;      // [nlev], [nk], [nlev]: scratch of the projection prologue
  int* live_k = kend + nlev;
  int* queue = live_k + nk;
  __shared__ int s_start;
  __shared__ uint32_t s_epoch;

  const size_t b = (size_t)c * nlev, bh = (size_t)c * (nlev + 1);
  const double zs = ld<T>(a.g.Zghalf, bh + nlev);
  const spc_gcm_tend& o = a.o;
  const size_t pfs = (size_t)ncol * nk;
  if (a.sync && threadIdx.x == 0) s_epoch = a.sync[SPC_SYNC_EPOCH];   // only this launch's last CTA changes it, at the very end

  // Stage every profile the column needs into shared memory first: all these global loads are independent and in flight
  // together (one memory latency), and the level loop further down runs out of shared memory only; the projection
  // prologue in between has its own chain of dependent loads (counts -> slab ranges -> mask rows).
  for (int l = threadIdx.x; l < nlev; l += blockDim.x) {
    ZfA[nlev - 1 - l] = (ld<T>(a.g.Zgfull, b + l) - zs) / grav;         // les.gcm_Zf, spcpl.py:198,390
    PfA[nlev - 1 - l] = ld<T>(a.g.Pfull, b + l);
    gT[l] = ld<T>(a.g.T, b + l);
    gSH[l] = ld<T>(a.g.SH, b + l);
    gQL[l] = ld<T>(a.g.QL, b + l);
    gQI[l] = ld<T>(a.g.QI, b + l);
    gU[l] = ld<T>(a.g.U, b + l);
    gV[l] = ld<T>(a.g.V, b + l);
    gA[l] = ld<T>(a.g.A, b + l);
  }
  if (a.conservative)
    for (int l = threadIdx.x; l <= nlev; l += blockDim.x) ZhD[l] = (ld<T>(a.g.Zghalf, bh + l) - zs) / grav;
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    const size_t i = (size_t)c * nk + k;
    zf[k] = __ldg(a.zf + k);
    const double ql = __ldg(a.les.prof + SPC_QL * pfs + i);
    const double qli = ld<T>(a.les.QL_ice, i);
    t_d[k] = ld<T>(a.les.T, i);
    qt_d[k] = __ldg(a.les.prof + SPC_QT * pfs + i);
    ql_d[k] = ql;
    qlw_d[k] = ql - qli;                                                // spcpl.py:402
    qli_d[k] = qli;
    u_d[k] = __ldg(a.les.prof + SPC_U * pfs + i);
    v_d[k] = __ldg(a.les.prof + SPC_V * pfs + i);
    if (a.conservative) {
      rho[k] = ld<T>(a.les.Rhobf, i);
      zh_s[k] = __ldg(a.zh + k);
    }
  }

  // cloud fraction source: given A | counts projected here from the KJI mask | counts written to o.cntslab by the
  // IJK projection kernel launched just before | none
  const bool fused = a.project_mw > 0;
  const bool from_global = !fused && (a.les.A == nullptr) && a.les.mask && o.cntslab;
  if (fused) {
    const int nq = cloud_slab_queue<T>(a.les.slab_idx, a.les.cnt, nk, nlev, c, kend, live_k, queue, o.cntslab, (T*)nullptr, cs_sh);
    const bool vec = (a.project_mw & 3) == 0 && (reinterpret_cast<uintptr_t>(a.les.mask) & 15) == 0;
    project_kji_queued<T>(a.les.mask + (size_t)c * nk * a.project_mw, vec, a.project_mw, nq, queue, kend, live_k, nlev, c,
                          1.0, o.cntslab, (T*)nullptr, cs_sh);
  }
  __syncthreads();

  // diagnostic temperature on LES levels, spcpl.py:408-409
  if (o.t || o.bracket_pf) {
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
      const size_t i = (size_t)c * nk + k;
      const int j = upper_bound(ZfA, nlev, zf[k]) - 1;
      if (o.bracket_pf) o.bracket_pf[i] = j;
      if (o.t) {
        const double pf = interp_at(ZfA, PfA, nlev, zf[k], j);
        const double thl_k = __ldg(a.les.prof + SPC_THL * pfs + i);
        st<T>(o.t, i, thl_k * pow(pf / pref0, kExp) + rlv * ql_d[k] / cp);
      }
    }
  }

  if (threadIdx.x == 0) {
    // start_index = searchsorted(-Zf, -h[-1]) (spcpl.py:498): GCM levels strictly above the LES top.
    // -Zf ascending <=> ZfA descending index; count of Zf > h_top = nlev - upper_bound(ZfA, h_top)
    s_start = nlev - upper_bound(ZfA, nlev, zf[nk - 1]);
    if (o.start_index) o.start_index[c] = s_start;
  }
  __syncthreads();

  const double npts = (double)a.les.nx * (double)a.les.ny;
  const double zh_top = a.conservative ? zh_s[nk - 1] : 0.0;
  const int tb = a.sync ? (int)(s_epoch % (uint32_t)a.n_bufs) * a.n_peers : 0;   // buffer set of this launch
  for (int l = threadIdx.x; l < nlev; l += blockDim.x) {
    const size_t i = b + l;
    const int r = nlev - 1 - l;                  // ascending slab index of GCM level l
    double A_d;                                  // profile["A"][::-1], spcpl.py:404
    if (a.les.A) A_d = ld<T>(a.les.A, b + r);
    else if (fused) A_d = (double)cs_sh[r] / npts;
    else if (from_global) A_d = (double)o.cntslab[b + r] / npts;   // cntslab is in ascending slab order
    else A_d = 0.0;
    st<T>(o.A_d, i, A_d);
    const double x = ZfA[r];                     // Zf[l]
    double val[7];                               // t, qt, ql, ql_water, ql_ice, u, v on GCM level l
    if (!a.conservative) {                       // spcpl.py:468-477
      const int j = upper_bound(zf, nk, x) - 1;
      if (o.bracket) o.bracket[i] = j;
      val[0] = interp_at(zf, t_d, nk, x, j);
      val[1] = interp_at(zf, qt_d, nk, x, j);
      val[2] = interp_at(zf, ql_d, nk, x, j);
      val[3] = interp_at(zf, qlw_d, nk, x, j);
      val[4] = interp_at(zf, qli_d, nk, x, j);
      val[5] = interp_at(zf, u_d, nk, x, j);
      val[6] = interp_at(zf, v_d, nk, x, j);
    } else {                                     // spcpl.py:479-488 -> sputils.interp_c (sputils.py:173-189)
#pragma unroll
      for (int n = 0; n < 7; ++n) val[n] = 0.0;
      if (o.bracket) o.bracket[i] = -2;
      if (ZhD[l] < zh_top) {                     // sputils.py:187
        const double* const q[7] = {t_d, qt_d, ql_d, qlw_d, qli_d, u_d, v_d};
        integral_w_multi<7>(ZhD[l + 1], ZhD[l], zh_s, nk, q, rho, val);
      }
    }
    const double ft = a.dt;                      // spcpl.py:427
    double f[SPC_NTEND];
    f[SPC_F_T] = a.factor * (val[0] - gT[l]) / ft;               // spcpl.py:518
    f[SPC_F_SH] = a.factor * ((val[1] - val[2]) - gSH[l]) / ft;  // spcpl.py:519
    f[SPC_F_QL] = a.factor * (val[3] - gQL[l]) / ft;             // spcpl.py:520
    f[SPC_F_QI] = a.factor * (val[4] - gQI[l]) / ft;             // spcpl.py:521
    f[SPC_F_U] = a.factor * (val[5] - gU[l]) / ft;               // spcpl.py:524
    f[SPC_F_V] = a.factor * (val[6] - gV[l]) / ft;               // spcpl.py:525
    f[SPC_F_A] = a.factor * (A_d - gA[l]) / ft;                  // spcpl.py:526
    const bool above = l < s_start;              // spcpl.py:527-533: f[0:start_index] *= 0
#pragma unroll
    for (int n = 0; n < SPC_NTEND; ++n) {
      const double v = above ? f[n] * 0.0 : f[n];
      const size_t e = ((size_t)c * SPC_NTEND + n) * nlev + l;
      st<T>(o.tend, e, v);
      for (int p = 0; p < a.n_peers; ++p) static_cast<T*>(a.peers[tb + p])[a.peer_off + e] = (T)v;   // NVLink / PCIe stores
    }
  }

  if (a.sync == nullptr) return;
  // ---- completion protocol (include/spcpl_b200.h, spc_gcm_tend.sync) ----
  __threadfence_system();                          // this thread's stores are visible to peers / the host
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  uint32_t last = 0;
  if (lane == 0) last = (atomicAdd(a.sync + SPC_SYNC_DONE, 1u) == gridDim.x - 1) ? 1u : 0u;
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence_system();                          // acquire: every CTA's fenced stores precede the flags below
  const uint32_t next = s_epoch + 1u;
  if (lane < a.n_signal) st_release_sys(a.signal[lane] + SPC_SYNC_FLAG0 + a.sync_slot, next);
  if (lane < a.n_wait) {
    const uint32_t* flag = a.sync + SPC_SYNC_FLAG0 + lane;
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(flag) - next) < 0) {
      __nanosleep(64);
      if (global_ns() - t0 > 20000000000ull) {     // a peer never arrived: record it and let the launch end
        atomicExch(a.sync + SPC_SYNC_ERROR, 1u + (uint32_t)lane);
        break;
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    a.sync[SPC_SYNC_DONE] = 0u;
    __threadfence();
    a.sync[SPC_SYNC_EPOCH] = next;
  }
}

// ------------------------------------------------------------------- sputils batch helpers
template <typename T>
__global__ void interp_kernel(const T* x, int x_batched, const T* xp, const T* fp, int nx, int np, T* out, int32_t* br) {
  extern __shared__ __align__(16) double sm[];
  double* sxp = sm;
  double* sfp = sm + np;
  const int bidx = blockIdx.x;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    sxp[i] = (double)xp[(size_t)bidx * np + i];
    sfp[i] = (double)fp[(size_t)bidx * np + i];
  }
  __syncthreads();
  const T* xr = x + (x_batched ? (size_t)bidx * nx : 0);
  for (int i = threadIdx.x; i < nx; i += blockDim.x) {
    const double xv = (double)xr[i];
    const int j = upper_bound(sxp, np, xv) - 1;
    if (br) br[(size_t)bidx * nx + i] = j;
    if (out) out[(size_t)bidx * nx + i] = (T)interp_at(sxp, sfp, np, xv, j);
  }
}

template <typename T>
__global__ void searchsorted_kernel(const T* a, const T* v, int v_batched, int na, int nv, int right, int32_t* out) {
  extern __shared__ __align__(16) double sm[];
  const int bidx = blockIdx.x;
  for (int i = threadIdx.x; i < na; i += blockDim.x) sm[i] = (double)a[(size_t)bidx * na + i];
  __syncthreads();
  const T* vr = v + (v_batched ? (size_t)bidx * nv : 0);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const double x = (double)vr[i];
    out[(size_t)bidx * nv + i] = right ? upper_bound(sm, na, x) : lower_bound(sm, na, x);
  }
}

// sputils.integral without weights (sputils.py:141-148)
__device__ double integral_plain(double a, double b, const double* z, int n, const double* q) {
  int ia = 0;
  while (ia + 1 < n - 1 && z[ia + 1] < a) ++ia;
  int ib = ia;
  while (ib + 1 < n - 1 && z[ib + 1] < b) ++ib;
  double S = 0.0;
  for (int i = ia; i <= ib; ++i) S += q[i] * (z[i + 1] - z[i]);        // sputils.py:142
  return S - q[ia] * (a - z[ia]) - q[ib] * (z[ib + 1] - b);             // sputils.py:145-148
}

// sputils.interp_c / interp_rho / integral over a batch of columns: one block per column, one thread per layer,
// the column's cell values staged in shared memory as float64.
template <typename T>
__global__ void __launch_bounds__(kThreads) interp_c_kernel(const T* Zh, const double* z, int nz, const T* q, const T* w, int nq,
                                                            int nlev, int mode, T* out) {
  extern __shared__ __align__(16) double sm[];
  double* zs = sm;             // [nz]
  double* qs = zs + nz;        // [nz-1]
  double* ws = qs + (nz - 1);  // [nz-1]
  const int bidx = blockIdx.x;
  for (int i = threadIdx.x; i < nz; i += blockDim.x) zs[i] = __ldg(z + i);
  for (int i = threadIdx.x; i < nz - 1; i += blockDim.x) {
    qs[i] = q ? (double)q[(size_t)bidx * nq + i] : 0.0;
    ws[i] = w ? (double)w[(size_t)bidx * nq + i] : 1.0;
  }
  __syncthreads();
  const T* Z = Zh + (size_t)bidx * (nlev + 1);
  for (int l = threadIdx.x; l < nlev; l += blockDim.x) {
    const double hi = (double)Z[l], lo = (double)Z[l + 1];
    double r = 0.0;
    if (mode == SPC_INT_C) {
      if (hi < zs[nz - 1]) r = integral_w(lo, hi, zs, nz, qs, ws);                    // sputils.py:187-188
    } else if (mode == SPC_INT_RHO) {
      if (hi < zs[nz - 1]) r = integral_plain(lo, hi, zs, nz, ws) / (hi - lo);         // sputils.py:195-196
    } else if (mode == SPC_INT_PLAIN) {
      r = integral_plain(lo, hi, zs, nz, qs);
    } else {
      r = integral_w(lo, hi, zs, nz, qs, ws);
    }
    out[(size_t)bidx * nlev + l] = (T)r;
  }
}

template <typename T>
__global__ void exner_kernel(const T* p, size_t n, double e, T* out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (T)pow((double)p[i] / pref0, e);
}

int check_gcm(const spc_gcm_cols* g, int couple_surface, const char* who) {
  SPC_REQUIRE(g != nullptr, SPC_ERR_ARG, "%s: gcm is NULL", who);
  SPC_REQUIRE(g->ncol >= 0 && g->nlev >= 2, SPC_ERR_ARG, "%s: bad shape ncol=%d nlev=%d", who, g->ncol, g->nlev);
  SPC_REQUIRE(g->dtype == SPC_F32 || g->dtype == SPC_F64, SPC_ERR_ARG, "%s: bad dtype %d", who, g->dtype);
  if (g->ncol == 0) return SPC_OK;
  SPC_REQUIRE(g->U && g->V && g->T && g->SH && g->QL && g->QI && g->Pfull && g->A && g->Zgfull && g->Phalf && g->Zghalf,
              SPC_ERR_ARG, "%s: a GCM profile pointer is NULL", who);
  if (couple_surface)
    SPC_REQUIRE(g->Z0M && g->Z0H && g->QLflux && g->QIflux && g->SHflux && g->TSflux, SPC_ERR_ARG,
                "%s: couple_surface set but a surface field pointer is NULL", who);
  return SPC_OK;
}

}  // namespace

extern "C" {

#ifdef SPC_TUNING
// Tuning hook of libspcpl_b200_tune.so (tools/step_probe.py), not part of the ABI: threads per column CTA.
int spc_tune_profiles(spc_handle h, int k2_threads, int k3_threads, int proj_threads) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  h->proj_threads = (proj_threads >= 32 && proj_threads <= 256) ? (proj_threads & ~31) : 0;
  h->k2_threads = (k2_threads >= 32 && k2_threads <= kThreads) ? (k2_threads & ~31) : 0;
  h->k3_threads = (k3_threads >= 32 && k3_threads <= kThreads) ? (k3_threads & ~31) : 0;
  return SPC_OK;
}
#endif

int spc_gcm_to_les(spc_handle h, const spc_gcm_cols* gcm, const double* zf, const double* zh, int nk,
                   const double* les_prof, const void* ps_les, double dt, double factor, int couple_surface,
                   const spc_les_forcing* out, void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  rc = check_gcm(gcm, couple_surface, "spc_gcm_to_les");
  if (rc) return rc;
  SPC_REQUIRE(zf && out && nk >= 1, SPC_ERR_ARG, "spc_gcm_to_les: zf/out NULL or nk < 1");
  SPC_REQUIRE(dt != 0.0, SPC_ERR_ARG, "spc_gcm_to_les: dt must be non-zero");
  SPC_REQUIRE(!(out->f_ps && !ps_les), SPC_ERR_ARG, "spc_gcm_to_les: f_ps requested but ps_les is NULL");
  SPC_REQUIRE(!(out->slab_idx && !zh), SPC_ERR_ARG, "spc_gcm_to_les: slab_idx requested but zh is NULL");
  SPC_REQUIRE(!((out->f_u || out->f_v || out->f_thl || out->f_qt || out->f_ql) && !les_prof), SPC_ERR_ARG,
              "spc_gcm_to_les: forcings requested but les_prof is NULL");
  if (gcm->ncol == 0) return SPC_OK;
  const size_t smem = ((size_t)6 * gcm->nlev + nk) * sizeof(double);
  SPC_REQUIRE(smem <= 48 * 1024, SPC_ERR_UNSUPPORTED, "spc_gcm_to_les: nlev=%d, nk=%d too large", gcm->nlev, nk);
  spc::DeviceGuard guard(h->device);
  K2Args a;
  a.g = to_ptrs(gcm);
  a.zf = zf; a.zh = zh; a.les_prof = les_prof; a.ps_les = ps_les;
  a.o = *out;
  a.dt = dt; a.factor = factor; a.nk = nk; a.couple_surface = couple_surface;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = column_threads(h->k2_threads, std::max(gcm->nlev + 1, nk));
  const bool one_wave = gcm->ncol <= 6 * h->num_sms;
  if (gcm->dtype == SPC_F32) {
    if (one_wave) gcm_to_les_kernel<float, true><<<gcm->ncol, threads, smem, st>>>(a);
    else gcm_to_les_kernel<float, false><<<gcm->ncol, threads, smem, st>>>(a);
  } else {
    if (one_wave) gcm_to_les_kernel<double, true><<<gcm->ncol, threads, smem, st>>>(a);
    else gcm_to_les_kernel<double, false><<<gcm->ncol, threads, smem, st>>>(a);
  }
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

int spc_les_to_gcm(spc_handle h, const spc_gcm_cols* gcm, const double* zf, const double* zh, int nk,
                   const spc_les_prof* les, double dt, double factor, int conservative, const spc_gcm_tend* out,
                   void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  rc = check_gcm(gcm, 0, "spc_les_to_gcm");
  if (rc) return rc;
  SPC_REQUIRE(zf && les && out && nk >= 1, SPC_ERR_ARG, "spc_les_to_gcm: zf/les/out NULL or nk < 1");
  SPC_REQUIRE(dt != 0.0, SPC_ERR_ARG, "spc_les_to_gcm: dt must be non-zero");
  SPC_REQUIRE(out->tend != nullptr, SPC_ERR_ARG, "spc_les_to_gcm: out->tend is NULL");
  const int n_peers = out->tend_peers ? out->n_peers : 0;
  const int n_bufs = out->sync ? out->n_bufs : 1;
  SPC_REQUIRE(n_peers >= 0 && n_peers <= SPC_MAX_PEERS && out->peer_col0 >= 0, SPC_ERR_ARG,
              "spc_les_to_gcm: n_peers=%d / peer_col0=%d out of range", out->n_peers, out->peer_col0);
  SPC_REQUIRE(n_bufs == 1 || n_bufs == 2, SPC_ERR_ARG, "spc_les_to_gcm: n_bufs=%d (1 or 2)", out->n_bufs);
  for (int p = 0; p < n_peers * n_bufs; ++p)
    SPC_REQUIRE(out->tend_peers[p] != nullptr, SPC_ERR_ARG, "spc_les_to_gcm: tend_peers[%d] is NULL", p);
  if (out->sync) {
    SPC_REQUIRE(out->n_signal >= 0 && out->n_signal <= SPC_MAX_PEERS + 1 && (out->n_signal == 0 || out->signal), SPC_ERR_ARG,
                "spc_les_to_gcm: n_signal=%d out of range or signal is NULL", out->n_signal);
    SPC_REQUIRE(out->n_wait >= 0 && out->n_wait <= SPC_SYNC_MAX_SLOTS && out->sync_slot >= 0 && out->sync_slot < SPC_SYNC_MAX_SLOTS,
                SPC_ERR_ARG, "spc_les_to_gcm: n_wait=%d / sync_slot=%d out of range", out->n_wait, out->sync_slot);
    for (int p = 0; p < out->n_signal; ++p)
      SPC_REQUIRE(out->signal[p] != nullptr, SPC_ERR_ARG, "spc_les_to_gcm: signal[%d] is NULL", p);
  }
  SPC_REQUIRE(les->prof && les->QL_ice && les->T, SPC_ERR_ARG, "spc_les_to_gcm: a LES profile pointer is NULL");
  SPC_REQUIRE(!(conservative && (!les->Rhobf || !zh)), SPC_ERR_ARG,
              "spc_les_to_gcm: conservative coarsening needs Rhobf and zh");
  const bool project = !les->A && les->mask;
  size_t mask_words = 0;
  if (project) {
    SPC_REQUIRE(les->slab_idx != nullptr, SPC_ERR_ARG, "spc_les_to_gcm: mask given without slab_idx");
    SPC_REQUIRE(les->nx > 0 && les->ny > 0, SPC_ERR_ARG, "spc_les_to_gcm: mask given without nx, ny");
    mask_words = spc_mask_words_per_column(les->vol_dtype, les->layout, les->nx, les->ny, nk);
    SPC_REQUIRE(mask_words > 0, SPC_ERR_UNSUPPORTED, "spc_les_to_gcm: no cloud mask format for this layout/shape");
    SPC_REQUIRE(les->layout == SPC_LAYOUT_KJI || out->cntslab != nullptr, SPC_ERR_ARG,
                "spc_les_to_gcm: out->cntslab is required when the cloud fraction comes from an IJK-layout mask");
  }
  SPC_REQUIRE(!(out->sync && gcm->ncol == 0), SPC_ERR_ARG, "spc_les_to_gcm: the completion protocol needs at least one column per rank");
  if (gcm->ncol == 0) return SPC_OK;
  // doubles: ZfA, PfA [nlev] | zf + 7 profiles + rho [9 nk] | ZhD [nlev+1] | zh [nk] | 7 GCM profiles [7 nlev]
  // ints: cs, kend, queue [nlev], live_k [nk]
  const size_t smem = ((size_t)10 * gcm->nlev + 1 + (size_t)10 * nk) * sizeof(double) + ((size_t)3 * gcm->nlev + nk) * sizeof(int);
  SPC_REQUIRE(smem <= 48 * 1024, SPC_ERR_UNSUPPORTED, "spc_les_to_gcm: nlev=%d, nk=%d too large", gcm->nlev, nk);
  spc::DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  K3Args a;
  a.project_mw = 0;
  if (project && les->layout == SPC_LAYOUT_KJI) {
    a.project_mw = (int)(mask_words / nk);       // the projection is K3's own prologue
  } else if (project) {                          // IJK mask: its projection kernel feeds K3 through out->cntslab
    rc = launch_cloud_projection<float>(h, les->mask, les->slab_idx, les->cnt, les->vol_dtype, les->layout, les->nx, les->ny,
                                        nk, gcm->ncol, gcm->nlev, out->cntslab, nullptr, st);
    if (rc) return rc;
  }
  a.g = to_ptrs(gcm);
  a.zf = zf; a.zh = zh; a.les = *les; a.o = *out;
  a.dt = dt; a.factor = factor; a.nk = nk; a.conservative = conservative;
  a.n_peers = n_peers;
  a.n_bufs = n_bufs;
  for (int p = 0; p < n_peers * n_bufs; ++p) a.peers[p] = out->tend_peers[p];
  a.peer_off = (size_t)out->peer_col0 * SPC_NTEND * gcm->nlev;
  a.sync = out->sync;
  a.n_signal = out->sync ? out->n_signal : 0;
  a.n_wait = out->sync ? out->n_wait : 0;
  a.sync_slot = out->sync_slot;
  for (int p = 0; p < a.n_signal; ++p) a.signal[p] = out->signal[p];
  const int threads = column_threads(h->k3_threads, std::max(gcm->nlev + 1, nk));
  if (gcm->dtype == SPC_F32) les_to_gcm_kernel<float><<<gcm->ncol, threads, smem, st>>>(a);
  else les_to_gcm_kernel<double><<<gcm->ncol, threads, smem, st>>>(a);
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

int spc_cloud_fraction(spc_handle h, const uint32_t* mask, const int32_t* slab_idx, const int32_t* cnt, int vol_dtype,
                       int layout, int nx, int ny, int nk, int ncol, int nlev, int out_dtype, int32_t* cntslab, void* A,
                       void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(ncol >= 0 && nlev >= 1 && nk >= 1 && nx > 0 && ny > 0, SPC_ERR_ARG, "spc_cloud_fraction: bad shape");
  SPC_REQUIRE(out_dtype == SPC_F32 || out_dtype == SPC_F64, SPC_ERR_ARG, "spc_cloud_fraction: bad out_dtype %d", out_dtype);
  if (ncol == 0) return SPC_OK;
  SPC_REQUIRE(mask && slab_idx && (cntslab || A), SPC_ERR_ARG, "spc_cloud_fraction: NULL pointer");
  spc::DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == SPC_F32)
    return launch_cloud_projection<float>(h, mask, slab_idx, cnt, vol_dtype, layout, nx, ny, nk, ncol, nlev, cntslab, (float*)A, st);
  return launch_cloud_projection<double>(h, mask, slab_idx, cnt, vol_dtype, layout, nx, ny, nk, ncol, nlev, cntslab, (double*)A, st);
}

int spc_interp(spc_handle h, int dtype, const void* x, int x_batched, const void* xp, const void* fp, int nb, int nx,
               int np, void* out, int32_t* bracket, void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(dtype == SPC_F32 || dtype == SPC_F64, SPC_ERR_ARG, "spc_interp: bad dtype %d", dtype);
  SPC_REQUIRE(nb >= 0 && nx >= 0 && np >= 1, SPC_ERR_ARG, "spc_interp: bad shape nb=%d nx=%d np=%d", nb, nx, np);
  if (nb == 0 || nx == 0) return SPC_OK;
  SPC_REQUIRE(x && xp && fp && (out || bracket), SPC_ERR_ARG, "spc_interp: NULL pointer");
  const size_t smem = (size_t)2 * np * sizeof(double);
  SPC_REQUIRE(smem <= 48 * 1024, SPC_ERR_UNSUPPORTED, "spc_interp: np=%d too large", np);
  spc::DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SPC_F32)
    interp_kernel<float><<<nb, kThreads, smem, st>>>((const float*)x, x_batched, (const float*)xp, (const float*)fp, nx, np,
                                                    (float*)out, bracket);
  else
    interp_kernel<double><<<nb, kThreads, smem, st>>>((const double*)x, x_batched, (const double*)xp, (const double*)fp, nx,
                                                     np, (double*)out, bracket);
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

int spc_searchsorted(spc_handle h, int dtype, const void* a, const void* v, int v_batched, int nb, int na, int nv,
                     int side_right, int32_t* out, void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(dtype == SPC_F32 || dtype == SPC_F64, SPC_ERR_ARG, "spc_searchsorted: bad dtype %d", dtype);
  SPC_REQUIRE(nb >= 0 && na >= 0 && nv >= 0, SPC_ERR_ARG, "spc_searchsorted: bad shape");
  if (nb == 0 || nv == 0) return SPC_OK;
  SPC_REQUIRE(v && out && (a || na == 0), SPC_ERR_ARG, "spc_searchsorted: NULL pointer");
  const size_t smem = (size_t)std::max(na, 1) * sizeof(double);
  SPC_REQUIRE(smem <= 48 * 1024, SPC_ERR_UNSUPPORTED, "spc_searchsorted: na=%d too large", na);
  spc::DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SPC_F32)
    searchsorted_kernel<float><<<nb, kThreads, smem, st>>>((const float*)a, (const float*)v, v_batched, na, nv, side_right, out);
  else
    searchsorted_kernel<double><<<nb, kThreads, smem, st>>>((const double*)a, (const double*)v, v_batched, na, nv, side_right,
                                                           out);
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

int spc_interp_c(spc_handle h, int dtype, const void* Zh, const double* z, int nz, const void* q, const void* w, int nq,
                 int nb, int nlev, int mode, void* out, void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(dtype == SPC_F32 || dtype == SPC_F64, SPC_ERR_ARG, "spc_interp_c: bad dtype %d", dtype);
  SPC_REQUIRE(mode >= SPC_INT_C && mode <= SPC_INT_WEIGHTED, SPC_ERR_ARG, "spc_interp_c: bad mode %d", mode);
  SPC_REQUIRE(nb >= 0 && nlev >= 0 && nz >= 2 && nq >= nz - 1, SPC_ERR_ARG, "spc_interp_c: bad shape nb=%d nlev=%d nz=%d nq=%d",
              nb, nlev, nz, nq);
  if (nb == 0 || nlev == 0) return SPC_OK;
  SPC_REQUIRE(Zh && z && out, SPC_ERR_ARG, "spc_interp_c: NULL pointer");
  SPC_REQUIRE(mode == SPC_INT_RHO || q, SPC_ERR_ARG, "spc_interp_c: q is NULL");
  SPC_REQUIRE(mode == SPC_INT_PLAIN || w, SPC_ERR_ARG, "spc_interp_c: this mode needs the weights w");
  const size_t smem = ((size_t)nz + 2 * (size_t)(nz - 1)) * sizeof(double);
  SPC_REQUIRE(smem <= 48 * 1024, SPC_ERR_UNSUPPORTED, "spc_interp_c: nz=%d too large", nz);
  spc::DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = column_threads(0, nlev);
  if (dtype == SPC_F32)
    interp_c_kernel<float><<<nb, threads, smem, st>>>((const float*)Zh, z, nz, (const float*)q, (const float*)w, nq, nlev, mode,
                                                      (float*)out);
  else
    interp_c_kernel<double><<<nb, threads, smem, st>>>((const double*)Zh, z, nz, (const double*)q, (const double*)w, nq, nlev,
                                                       mode, (double*)out);
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

int spc_exner(spc_handle h, int dtype, const void* p, size_t n, int inverse, void* out, void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(dtype == SPC_F32 || dtype == SPC_F64, SPC_ERR_ARG, "spc_exner: bad dtype %d", dtype);
  if (n == 0) return SPC_OK;
  SPC_REQUIRE(p && out, SPC_ERR_ARG, "spc_exner: NULL pointer");
  spc::DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)h->num_sms * 8);
  const double e = inverse ? kIExp : kExp;
  if (dtype == SPC_F32) exner_kernel<float><<<grid, kThreads, 0, st>>>((const float*)p, n, e, (float*)out);
  else exner_kernel<double><<<grid, kThreads, 0, st>>>((const double*)p, n, e, (double*)out);
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}

}  // extern "C"
