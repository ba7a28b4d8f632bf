// spc_variability_nudge: qt-variability nudging (spcpl.variability_nudge, splib/spcpl.py:613-744).
//
// Every (column, level) slab is independent: one thread block per slab keeps the slab's qt and qsat
// in shared memory, and all threads run the same Brent iteration (scipy.optimize.brentq restated,
// spcpl.py:672,708) in lock-step; each function evaluation is a fixed-order block reduction over the
// slab (a "slab reduce inside a root finder"). The volumes are read once and the
// nudged qt (and thl) written once. Compiled with --fmad=false like the other profile kernels.
#include <float.h>

#include "spc_common.cuh"

namespace {

using namespace spc;

constexpr int kThreads = 256;
constexpr double kXtol = 2e-12;                    // scipy/optimize/_zeros_py.py: _xtol
constexpr double kRtol = 4 * DBL_EPSILON;          // _rtol = 4 * eps
constexpr int kMaxIter = 100;                      // _iter
constexpr double kBetaMax = 5.0;                   // spcpl.py:654-655, 702-703

enum { ST_MULT = 1, ST_UNSAT = 2, ST_ADD = 4, ST_NOBRACKET = 8, ST_ADD_FAIL = 16, ST_NOCONV = 32 };

struct NudgeArgs {
  void* qt;
  const void* qsat;
  const void* qsat_prof;
  void* thl;
  const void* ql;
  const double* prof;
  const void* ql_ref;
  const void* presf;
  const double* R;
  double DT;
  double *beta, *alpha, *qt_std;
  int32_t* status;
  int constant_T, ncol, nk, nx, ny, S;
};

// fixed-order block sum; every thread returns the total
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // red[] may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) s += red[w];
  return s;
}

template <typename T, bool CACHE>
struct Slab {
  const T* q;    // qt of the slab (shared-memory copy when CACHE)
  const T* qs;   // qsat of the slab, or nullptr when qsat is a profile value
  double qs0;    // profile qsat
  int S;
  __device__ __forceinline__ double qt(int e) const { return (double)q[e]; }
  __device__ __forceinline__ double qsat(int e) const { return qs ? (double)qs[e] : qs0; }
};

// get_ql_diff(beta), spcpl.py:641-643
template <typename SL>
__device__ __forceinline__ double ql_diff(const SL& s, double beta, double qt_av, double ql_ref, double* red) {
  double acc = 0.0;
  for (int e = threadIdx.x; e < s.S; e += kThreads) acc += fmax(beta * (s.qt(e) - qt_av) + qt_av - s.qsat(e), 0.0);
  return block_sum(acc, red) / (double)s.S - ql_ref;
}
// get_ql_diff_additive(a), spcpl.py:648-651
template <typename SL>
__device__ __forceinline__ double ql_diff_add(const SL& s, double a, const double* R, double ql_ref, double* red) {
  double acc = 0.0;
  for (int e = threadIdx.x; e < s.S; e += kThreads) acc += fmax(s.qt(e) + (a * __ldg(R + e)) - s.qsat(e), 0.0);
  return block_sum(acc, red) / (double)s.S - ql_ref;
}

// scipy.optimize.brentq (Zeros/brentq.c), executed redundantly by every thread of the block with
// block-uniform function values. status: 0 ok, -1 no sign change, -2 not converged.
template <typename F>
__device__ double brentq(F f, double xa, double xb, int& status) {
  double xpre = xa, xcur = xb, xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
  double fpre = f(xpre), fcur = f(xcur);
  status = 0;
  if (fpre == 0) return xpre;
  if (fcur == 0) return xcur;
  if (signbit(fpre) == signbit(fcur)) {
    status = -1;
    return 0.0;
  }
  for (int i = 0; i < kMaxIter; ++i) {
    if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
      xblk = xpre;
      fblk = fpre;
      spre = scur = xcur - xpre;
    }
    if (fabs(fblk) < fabs(fcur)) {
      xpre = xcur; xcur = xblk; xblk = xpre;
      fpre = fcur; fcur = fblk; fblk = fpre;
    }
    const double delta = (kXtol + kRtol * fabs(xcur)) / 2;
    const double sbis = (xblk - xcur) / 2;
    if (fcur == 0 || fabs(sbis) < delta) return xcur;
    if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
      double stry;
      if (xpre == xblk) {
        stry = -fcur * (xcur - xpre) / (fcur - fpre);
      } else {
        const double dpre = (fpre - fcur) / (xpre - xcur);
        const double dblk = (fblk - fcur) / (xblk - xcur);
        stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
      }
      if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) {
        spre = scur;
        scur = stry;
      } else {
        spre = sbis;
        scur = sbis;
      }
    } else {
      spre = sbis;
      scur = sbis;
    }
    xpre = xcur;
    fpre = fcur;
    if (fabs(scur) > delta) xcur += scur;
    else xcur += (sbis > 0 ? delta : -delta);
    fcur = f(xcur);
  }
  status = -2;
  return xcur;
}

template <typename T, bool CACHE>
__global__ void __launch_bounds__(kThreads) nudge_kernel(const NudgeArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ double red[kThreads / 32];
  __shared__ double s_best[kThreads / 32];
  __shared__ int s_besti[kThreads / 32];
  const int S = a.S;
  const size_t slab = blockIdx.x;                 // = c*nk + k
  const int c = (int)(slab / a.nk);
  T* gq = static_cast<T*>(a.qt) + slab * S;
  const T* gqs = a.qsat ? static_cast<const T*>(a.qsat) + slab * S : nullptr;
  Slab<T, CACHE> s;
  s.S = S;
  s.qs0 = a.qsat_prof ? (double)static_cast<const T*>(a.qsat_prof)[slab] : 0.0;
  if constexpr (CACHE) {
    T* sq = reinterpret_cast<T*>(smem_raw);
    T* sqs = sq + S;
    for (int e = threadIdx.x; e < S; e += kThreads) {
      sq[e] = gq[e];
      if (gqs) sqs[e] = gqs[e];
    }
    __syncthreads();
    s.q = sq;
    s.qs = gqs ? sqs : nullptr;
  } else {
    s.q = gq;
    s.qs = gqs;
  }
  const size_t pfs = (size_t)a.ncol * a.nk;
  const double qt_av = __ldg(a.prof + SPC_QT * pfs + slab);       // les.get_profile("QT"), spcpl.py:630
  const double ql_av = __ldg(a.prof + SPC_QL * pfs + slab);       // les.get_profile("QL"), spcpl.py:629
  const double ql_ref = (double)static_cast<const T*>(a.ql_ref)[slab];
  const double* R = a.R ? a.R + (size_t)c * S : nullptr;

  double beta = 1.0;                                               // spcpl.py:657
  int st = 0;
  bool touched = true;                                             // false = `continue`, spcpl.py:693
  double add_a = 0.0;
  bool additive = false;
  if (ql_ref > 1e-9) {                                             // spcpl.py:661
    const double q_min = ql_diff(s, 0.0, qt_av, ql_ref, red);      // spcpl.py:663-664
    const double q_max = ql_diff(s, kBetaMax, qt_av, ql_ref, red);
    if (q_min > 0 || q_max < 0) {                                  // spcpl.py:665-669
      beta = kBetaMax;
      st |= ST_NOBRACKET;
    } else {
      int bs;
      beta = brentq([&](double b) { return ql_diff(s, b, qt_av, ql_ref, red); }, 0.0, kBetaMax, bs);   // :672
      st |= ST_MULT;
      if (bs == -2) st |= ST_NOCONV;                               // scipy would raise RuntimeError (maxiter) here
    }
  } else if (ql_av > ql_ref) {                                     // spcpl.py:675-691
    // argmax(qt - qsat) over the slab, first maximum in the reference's (i, j) C order
    double best = -DBL_MAX;
    int besti = 0x7fffffff;
    for (int e = threadIdx.x; e < S; e += kThreads) {
      const double d = s.qt(e) - s.qsat(e);
      const int pri = (e % a.nx) * a.ny + e / a.nx;
      if (d > best || (d == best && pri < besti)) {
        best = d;
        besti = pri;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) {
        best = ob;
        besti = oi;
      }
    }
    if ((threadIdx.x & 31) == 0) {
      s_best[threadIdx.x >> 5] = best;
      s_besti[threadIdx.x >> 5] = besti;
    }
    __syncthreads();
    best = s_best[0];
    besti = s_besti[0];
    for (int w = 1; w < kThreads / 32; ++w)
      if (s_best[w] > best || (s_best[w] == best && s_besti[w] < besti)) {
        best = s_best[w];
        besti = s_besti[w];
      }
    const int e = (besti % a.ny) * a.nx + besti / a.ny;            // back from (i*ny + j) to j*nx + i
    beta = (s.qsat(e) - qt_av) / (s.qt(e) - qt_av);                // spcpl.py:678
    st |= ST_UNSAT;
    if (beta < 0) beta = 1.0;                                      // spcpl.py:688-691
  } else {
    touched = false;
  }
  bool mult = touched;
  if (touched && beta >= kBetaMax) {                               // spcpl.py:698-717
    mult = false;
    if (ql_ref > ql_av) {
      if (R != nullptr) {
        int bs;
        add_a = brentq([&](double x) { return ql_diff_add(s, x, R, ql_ref, red); }, 0.0, kBetaMax, bs);   // :708
        if (bs == -1) st |= ST_ADD_FAIL;                           // scipy would raise ValueError here
        else {
          additive = true;
          st |= ST_ADD;
          if (bs == -2) st |= ST_NOCONV;
        }
      } else {
        st |= ST_ADD_FAIL;
      }
    }
    beta = 1.0;                                                    // spcpl.py:717
  }
  // apply: qt += (beta-1)(qt - qt_av)  or  qt += a*R   (spcpl.py:711-720), constant-T theta_l
  // correction (spcpl.py:721-728), and the statistics of the nudged field (spcpl.py:737-741)
  double dthl_fac = 0.0;
  if (a.constant_T && touched)
    dthl_fac = -rlv / (cp * pow((double)static_cast<const T*>(a.presf)[slab] / pref0, rd / cp));
  double sum = 0.0;
  for (int e = threadIdx.x; e < S; e += kThreads) {
    const double q = s.qt(e);
    double qn = q;
    if (mult) qn = q + (beta - 1) * (q - qt_av);
    else if (additive) qn = q + add_a * __ldg(R + e);
    sum += qn;
  }
  const double mean = block_sum(sum, red) / (double)S;
  double var = 0.0;
  for (int e = threadIdx.x; e < S; e += kThreads) {
    const double q = s.qt(e);
    double qn = q;
    if (mult) qn = q + (beta - 1) * (q - qt_av);
    else if (additive) qn = q + add_a * __ldg(R + e);
    const double d = qn - mean;
    var += d * d;
    if (mult || additive) gq[e] = (T)qn;
    if (a.constant_T && touched) {
      const double ql_target = fmax(qn - s.qsat(e), 0.0);
      const double dQL = ql_target - (double)static_cast<const T*>(a.ql)[slab * S + e];
      T* th = static_cast<T*>(a.thl) + slab * S + e;
      *th = (T)((double)*th + dthl_fac * dQL);
    }
  }
  var = block_sum(var, red) / (double)S;
  if (threadIdx.x == 0) {
    if (a.beta) a.beta[slab] = beta;
    if (a.alpha) a.alpha[slab] = log(beta) / a.DT;                 // spcpl.py:737
    if (a.qt_std) a.qt_std[slab] = sqrt(var);                      // spcpl.py:741
    if (a.status) a.status[slab] = st;
  }
}

}  // namespace

extern "C" int spc_variability_nudge(spc_handle h, const spc_nudge_io* io, int dtype, int ncol, int nx, int ny, int nk,
                                     double DT, int constant_T, double* beta, double* alpha, double* qt_std,
                                     int32_t* status, void* stream) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(io != nullptr, SPC_ERR_ARG, "spc_variability_nudge: io is NULL");
  SPC_REQUIRE(dtype == SPC_F32 || dtype == SPC_F64, SPC_ERR_ARG, "spc_variability_nudge: bad dtype %d", dtype);
  SPC_REQUIRE(ncol >= 0 && nx > 0 && ny > 0 && nk > 0, SPC_ERR_ARG, "spc_variability_nudge: bad shape");
  SPC_REQUIRE(DT != 0.0, SPC_ERR_ARG, "spc_variability_nudge: DT must be non-zero");
  if (ncol == 0) return SPC_OK;
  SPC_REQUIRE(io->qt && io->prof && io->ql_ref, SPC_ERR_ARG, "spc_variability_nudge: qt/prof/ql_ref is NULL");
  SPC_REQUIRE((io->qsat != nullptr) != (io->qsat_prof != nullptr), SPC_ERR_ARG,
              "spc_variability_nudge: give exactly one of qsat (volume) and qsat_prof (profile)");
  SPC_REQUIRE(!constant_T || (io->thl && io->ql && io->presf), SPC_ERR_ARG,
              "spc_variability_nudge: constant_T needs thl, ql and presf");
  const long long S = (long long)nx * ny;
  SPC_REQUIRE(S < (1ll << 30), SPC_ERR_UNSUPPORTED, "spc_variability_nudge: slab too large");
  spc::DeviceGuard guard(h->device);
  NudgeArgs a;
  a.qt = io->qt; a.qsat = io->qsat; a.qsat_prof = io->qsat_prof; a.thl = io->thl; a.ql = io->ql; a.prof = io->prof;
  a.ql_ref = io->ql_ref; a.presf = io->presf; a.R = io->R; a.DT = DT;
  a.beta = beta; a.alpha = alpha; a.qt_std = qt_std; a.status = status;
  a.constant_T = constant_T; a.ncol = ncol; a.nk = nk; a.nx = nx; a.ny = ny; a.S = (int)S;
  const size_t esize = dtype == SPC_F32 ? 4 : 8;
  const size_t smem = 2 * (size_t)S * esize;
  const bool cache = smem <= 96 * 1024;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((size_t)ncol * nk);
  if (dtype == SPC_F32) {
    if (cache) {
      SPC_CUDA(cudaFuncSetAttribute(nudge_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      nudge_kernel<float, true><<<grid, kThreads, smem, st>>>(a);
    } else {
      nudge_kernel<float, false><<<grid, kThreads, 0, st>>>(a);
    }
  } else {
    if (cache) {
      SPC_CUDA(cudaFuncSetAttribute(nudge_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      nudge_kernel<double, true><<<grid, kThreads, smem, st>>>(a);
    } else {
      nudge_kernel<double, false><<<grid, kThreads, 0, st>>>(a);
    }
  }
  SPC_CUDA(cudaGetLastError());
  return SPC_OK;
}
