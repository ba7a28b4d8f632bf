/* spcpl_b200.h — C ABI of the B200-native per-column GCM<->LES coupling step.
 *
 * Drop-in boundary for ONE path of CloudResolvingClimateModeling/sp-coupler: the per-column
 * coupling math of splib/spcpl.py + splib/sputils.py, batched over [ncol, ...] device arrays.
 * The reference has no FFI for this path (it is pure Python over AMUSE RPC stubs), so each entry
 * point below cites the reference FUNCTION it replaces; the binding a maintainer adds on the
 * reference side is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller; kernels never allocate; outputs are
 *     written into caller-provided buffers. Optional outputs may be NULL.
 *   - `dtype` (SPC_F32 | SPC_F64) is the storage type of LES volumes, GCM input profiles, the
 *     LES-internal pass-through profiles and all floating-point outputs. Slab means travel
 *     between entry points as float64 ([5][ncol][nk], order THL,QT,QL,U,V); zf/zh are float64;
 *     counts and indices are int32. All arithmetic is done in float64 registers.
 *   - GCM arrays run top -> bottom ([ncol][nlev], half-level arrays [ncol][nlev+1] ending at the
 *     ground), LES arrays bottom -> top, as in the reference (spcpl.py:197,224-228).
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     returns 0 on success, <0 for an invalid argument/shape/alignment, >0 = cudaError_t.
 *     spc_last_error() returns a thread-local message for the last non-zero return.
 *   - Plain SI units throughout (the reference's AMUSE units are all SI-coherent).
 */
#ifndef SPCPL_B200_H
#define SPCPL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPC_ABI_VERSION 2

enum { SPC_F32 = 0, SPC_F64 = 1 };

/* Memory order of an LES volume. KJI: [ncol][nk][ny][nx], a horizontal slab is contiguous (DALES
 * native order). IJK: [ncol][nx][ny][nk], k fastest — the (itot,jtot,ktot) C-order view that OMUSE
 * hands to Python (spcpl.py:275,288,642). */
enum { SPC_LAYOUT_KJI = 0, SPC_LAYOUT_IJK = 1 };

enum {
  SPC_OK = 0,
  SPC_ERR_ARG = -1,         /* NULL/negative/inconsistent argument */
  SPC_ERR_ALIGN = -2,       /* a volume pointer is not 16-byte aligned */
  SPC_ERR_UNSUPPORTED = -3, /* shape too large for the on-chip profile buffers */
  SPC_ERR_HANDLE = -4
};

/* order of the five slab-averaged fields (spcpl.py:748-755) */
enum { SPC_THL = 0, SPC_QT = 1, SPC_QL = 2, SPC_U = 3, SPC_V = 4, SPC_NFIELDS = 5 };
/* order of the packed GCM tendency block (spcpl.py:518-526, 535-542) */
enum { SPC_F_T = 0, SPC_F_SH = 1, SPC_F_QL = 2, SPC_F_QI = 3, SPC_F_U = 4, SPC_F_V = 5, SPC_F_A = 6, SPC_NTEND = 7 };

typedef struct spc_ctx* spc_handle;

int spc_abi_version(void);
const char* spc_last_error(void);

/* One handle per device: caches SM count / shared-memory limits and opts kernels in to large
 * dynamic shared memory. No hidden per-step state. */
int spc_create(spc_handle* out, int device);
int spc_destroy(spc_handle h);

/* Exchange buffers in HOST memory that K3 may store into directly (spc_gcm_tend.tend_peers / signal): the
 * reference hands every tendency profile back to the host GCM (gcm.set_profile_tendency, spcpl.py:535-542).
 * spc_host_register pins [p, p+nbytes) - e.g. a /dev/shm mapping shared by the ranks of a node - and maps it for the
 * handle's device; spc_host_device_pointer returns the device-visible address of memory that is already pinned
 * (cudaHostAlloc, torch pin_memory). */
int spc_host_register(spc_handle h, void* p, size_t nbytes, void** dev_ptr);
int spc_host_unregister(spc_handle h, void* p);
int spc_host_device_pointer(spc_handle h, void* p, void** dev_ptr);

/* ---------------------------------------------------------------------------------------------
 * K1  slab_reduce — replaces the LES-side slab averages the reference requests in
 * spcpl.get_les_profiles (spcpl.py:747-759: get_profile_{THL,QT,QL,U,V}) and in the first-step
 * branch of set_les_forcings (spcpl.py:302-307), plus the above-threshold ql count behind
 * les.get_cloudfraction (spcpl.py:28,765).
 *
 *   vol[5]   THL,QT,QL,U,V volumes, each ncol*nk*ny*nx elements of `dtype` in `layout`
 *   prof     out float64 [5][ncol][nk]   mean over the nx*ny horizontal points
 *   cnt      out int32   [ncol][nk]      #{(i,j): (double)ql > ql_thresh}   (may be NULL)
 *   mask     out uint32  [ncol][spc_mask_words_per_column(...)]  per-cell cloud bitmask in an
 *            opaque layout (one bit per cell; differs between KJI and IJK) consumed by
 *            spc_les_to_gcm / spc_cloud_fraction for the projected cloud cover (may be NULL)
 */
size_t spc_mask_words_per_column(int dtype, int layout, int nx, int ny, int nk);

int spc_slab_reduce(spc_handle h, const void* const vol[5], int dtype, int layout,
                    int ncol, int nx, int ny, int nk, double ql_thresh,
                    double* prof, int32_t* cnt, uint32_t* mask, void* stream);

/* GCM state of the superparameterized columns: the reference's gcm_vars and surf_vars
 * (spcpl.py:32-33) as gathered by spcpl.gather_gcm_data (spcpl.py:55-86), struct-of-arrays. */
typedef struct {
  int ncol, nlev, dtype;
  const void *U, *V, *T, *SH, *QL, *QI, *Pfull, *A, *Zgfull; /* [ncol][nlev]   */
  const void *Phalf, *Zghalf;                                  /* [ncol][nlev+1] */
  /* surface fields [ncol]; only read when couple_surface != 0 */
  const void *Z0M, *Z0H, *QLflux, *QIflux, *SHflux, *TLflux, *TSflux;
} spc_gcm_cols;

/* ---------------------------------------------------------------------------------------------
 * K2  gcm_to_les — replaces spcpl.convert_profiles (spcpl.py:171-246), the forcing arithmetic of
 * spcpl.set_les_forcings (spcpl.py:328-333,347-348), spcpl.convert_surface_fluxes
 * (spcpl.py:136-167), sputils.iexner (sputils.py:33-34), sputils.interp (sputils.py:82-86) and
 * the cloud-fraction slab mapping sputils.searchsorted(zh, Zh, side="right")[:-1][::-1]
 * (spcpl.py:26,764).   f_x = factor * (x_gcm->les - <x>_les) / dt.
 */
typedef struct {
  void *f_u, *f_v, *f_thl, *f_qt, *f_ql; /* [ncol][nk] forcings (set_tendency_*)               */
  void *ql_ref;                          /* [ncol][nk] GCM QL on LES levels (set_ref_profile_QL) */
  void *u, *v, *thl, *qt;                /* [ncol][nk] optional: convert_profiles' return values */
  void *f_ps, *ps;                       /* [ncol] surface-pressure forcing, GCM surface pressure */
  void *z0m, *z0h, *wthl, *wqt;          /* [ncol] when couple_surface                           */
  void *Tv, *THL, *QT, *Zf;              /* [ncol][nlev]   optional diagnostics (spcpl.py:231-244) */
  void *Zh;                              /* [ncol][nlev+1] optional (les.gcm_Zh)                  */
  int32_t* bracket;                      /* [ncol][nk]   optional: upper_bound(Zf[::-1], zf)-1    */
  int32_t* slab_idx;                     /* [ncol][nlev] optional (needs zh): cloud slab mapping  */
} spc_les_forcing;

int spc_gcm_to_les(spc_handle h, const spc_gcm_cols* gcm, const double* zf, const double* zh, int nk,
                   const double* les_prof /* [5][ncol][nk] from spc_slab_reduce */,
                   const void* ps_les /* [ncol] LES surface pressure, dtype */,
                   double dt, double factor, int couple_surface,
                   const spc_les_forcing* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3  les_to_gcm — replaces spcpl.set_gcm_tendencies (spcpl.py:388-555) and the cloud fraction
 * of spcpl.get_cloud_fraction / get_les_profiles (spcpl.py:22-29,761-765), sputils.exner
 * (sputils.py:28-29), and, with conservative != 0, sputils.interp_c/integral (sputils.py:94-189).
 *   f_X = factor * (<x>_les->gcm - X) / dt, zeroed for levels above the LES top.
 */
typedef struct {
  const double* prof;      /* [5][ncol][nk] slab means from spc_slab_reduce                    */
  const void *QL_ice, *T;  /* [ncol][nk] LES-internal profiles (get_profile_QL_ice / _T)       */
  const void* Rhobf;       /* [ncol][nk] only read when conservative != 0                      */
  const void* A;           /* optional [ncol][nlev] cloud fraction, ascending slab order, dtype */
  const uint32_t* mask;    /* else: cloud mask from spc_slab_reduce + slab_idx below            */
  const int32_t* slab_idx; /* [ncol][nlev] from spc_gcm_to_les                                 */
  const int32_t* cnt;      /* optional [ncol][nk] counts from spc_slab_reduce: lets cloud-free
                              levels be skipped without reading their mask words               */
  int vol_dtype, layout, nx, ny; /* describe the volumes `mask` was built from                 */
} spc_les_prof;

typedef struct {
  void* tend;           /* packed [ncol][7][nlev]: f_T,f_SH,f_QL,f_QI,f_U,f_V,f_A              */
  void* t;              /* optional [ncol][nk]  diagnostic temperature (spcpl.py:408-409)       */
  void* A_d;            /* optional [ncol][nlev] LES cloud fraction in GCM order (spcpl.py:404) */
  int32_t* cntslab;     /* optional [ncol][nlev] projected cloudy-column count, ascending slabs (the cloud fraction K3
                           uses when it comes from les->mask). REQUIRED only for the IJK mask layout, where a separate
                           projection kernel hands it to K3; the KJI projection runs inside K3 itself. */
  int32_t* bracket;     /* optional [ncol][nlev] upper_bound(zf, Zf)-1                          */
  int32_t* bracket_pf;  /* optional [ncol][nk]   upper_bound(Zf[::-1], zf)-1                    */
  int32_t* start_index; /* optional [ncol]       searchsorted(-Zf, -zf[-1]) (spcpl.py:498)      */
  /* Fused gather / host exchange: besides `tend`, the kernel stores this rank's block straight into
   * n_peers remote buffers [ncol_total][7][nlev] at column offset peer_col0. A target is any address the
   * device can write: a peer GPU's buffer mapped over NVLink (symmetric memory), or PINNED HOST memory of the
   * process that owns the GCM (zero-copy stores over PCIe) - so the tendencies reach their consumer from K3's
   * own epilogue instead of through a separate collective or copy (reference analogue: the 7
   * gcm.set_profile_tendency calls per column, spcpl.py:535-542). tend_peers is a HOST array of
   * n_peers*n_bufs device-visible pointers (target p of buffer set b at [b*n_peers + p]); NULL / 0 disables it. */
  void* const* tend_peers;
  int n_peers;          /* <= SPC_MAX_PEERS */
  int peer_col0;
  /* Completion protocol of the launch (optional; sync == NULL disables it and n_bufs is taken as 1).
   *   sync      DEVICE memory of this rank, SPC_SYNC_WORDS uint32 words, zeroed once by the caller:
   *             [SPC_SYNC_EPOCH] launches completed so far, [SPC_SYNC_DONE] CTA counter (kernel-internal),
   *             [SPC_SYNC_ERROR] non-zero after a wait timed out (20 s), [SPC_SYNC_FLAG0 + s] flag of slot s.
   *   The launch with epoch e stores into buffer set (e % n_bufs). When its last CTA has finished and every store
   *   of the launch is visible system-wide, word SPC_SYNC_FLAG0 + sync_slot of each of the n_signal blocks in
   *   `signal` (HOST array of device-visible pointers: peer sync blocks, or a block in pinned host memory that the
   *   host polls) is set to e+1 with release semantics; then the kernel waits until the flags of slots
   *   [0, n_wait) of its OWN block have reached e+1 (a barrier over n_wait ranks when every rank signals every
   *   rank), and sets its epoch to e+1. Everything is inside the kernel: the step stays capturable in a CUDA graph. */
  int n_bufs;           /* 1 or 2 */
  uint32_t* sync;
  uint32_t* const* signal;
  int n_signal;         /* <= SPC_MAX_PEERS + 1 */
  int sync_slot;
  int n_wait;           /* <= SPC_SYNC_MAX_SLOTS */
} spc_gcm_tend;
#define SPC_MAX_PEERS 16
enum { SPC_SYNC_EPOCH = 0, SPC_SYNC_DONE = 1, SPC_SYNC_ERROR = 2, SPC_SYNC_FLAG0 = 8, SPC_SYNC_MAX_SLOTS = 32,
       SPC_SYNC_WORDS = 64 };

int spc_les_to_gcm(spc_handle h, const spc_gcm_cols* gcm, const double* zf, const double* zh, int nk,
                   const spc_les_prof* les, double dt, double factor, int conservative,
                   const spc_gcm_tend* out, void* stream);

/* les.get_cloudfraction(indices) (spcpl.py:28,765) on its own: projected cloudy-column count per
 * GCM slab from the spc_slab_reduce mask. slab_idx [ncol][nlev] ascending (spc_gcm_to_les);
 * cntslab int32 [ncol][nlev] and/or A = cntslab/(nx*ny) of out_dtype, both in ascending slab order
 * (the order get_cloudfraction returns; the coupler reverses it, spcpl.py:28,404). */
int spc_cloud_fraction(spc_handle h, const uint32_t* mask, const int32_t* slab_idx,
                       const int32_t* cnt /* optional [ncol][nk], as in spc_les_prof */,
                       int vol_dtype, int layout, int nx, int ny, int nk, int ncol, int nlev, int out_dtype,
                       int32_t* cntslab, void* A, void* stream);

/* ---------------------------------------------------------------------------------------------
 * sputils.interp / sputils.searchsorted (sputils.py:82-91) over a batch of rows.
 *   x  [nb][nx] (x_batched != 0) or [nx] shared by all rows; xp, fp [nb][np], xp increasing.
 *   out [nb][nx] = numpy.interp(x, xp, fp) (clamped);  bracket [nb][nx] = upper_bound(xp,x)-1.
 * spc_searchsorted: out int32 [nb][nv] = numpy.searchsorted(a[nb][na], v, side). */
int spc_interp(spc_handle h, int dtype, const void* x, int x_batched, const void* xp, const void* fp,
               int nb, int nx, int np, void* out, int32_t* bracket, void* stream);
int spc_searchsorted(spc_handle h, int dtype, const void* a, const void* v, int v_batched,
                     int nb, int na, int nv, int side_right, int32_t* out, void* stream);
/* sputils.exner / iexner (sputils.py:28-34): out[i] = (p[i]/pref0)^(+-rd/cp) */
int spc_exner(spc_handle h, int dtype, const void* p, size_t n, int inverse, void* out, void* stream);
/* sputils.integral / interp_c / interp_rho (sputils.py:94-197): integrals of a piecewise-constant profile
 * (q[nb][nq] on the cells [z[i], z[i+1]] of the ascending float64 edges z[nz], nq >= nz-1) over the layers
 * [Zh[i+1], Zh[i]] of the descending edges Zh[nb][nlev+1]; out[nb][nlev]. mode:
 *   SPC_INT_C       interp_c  : integral(q*w) / integral(w) where Zh[i] < z[nz-1], else 0 (sputils.py:173-189)
 *   SPC_INT_RHO     interp_rho: integral(w) / (Zh[i] - Zh[i+1])  where Zh[i] < z[nz-1], else 0 (:191-197; q unused)
 *   SPC_INT_PLAIN   integral(a=Zh[i+1], b=Zh[i], z, q)            (sputils.py:141-148, w == NULL)
 *   SPC_INT_WEIGHTED integral(a, b, z, q, w)                      (sputils.py:150-161)
 * The last two apply no range test: the caller keeps a and b inside [z[0], z[nz-1]] as the reference requires. */
enum { SPC_INT_C = 0, SPC_INT_RHO = 1, SPC_INT_PLAIN = 2, SPC_INT_WEIGHTED = 3 };
int spc_interp_c(spc_handle h, int dtype, const void* Zh, const double* z, int nz, const void* q, const void* w,
                 int nq, int nb, int nlev, int mode, void* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * spcpl.set_les_state (spcpl.py:274-294): broadcast a vertical profile to a volume with uniform
 * noise, vol[c,k,j,i] = prof[c,k] + amp*n - sub[c,k], n ~ U[-1,1) from Philox4x32-10 keyed by
 * (seed, stream_id) and counted by (element/4, col0+c); clamped at 0 when clamp0 (synthetic ql).
 * KJI layout. The reference draws from numpy's Mersenne Twister; only the distribution matches. */
int spc_set_les_state(spc_handle h, const double* prof, double amp, uint32_t stream_id, uint32_t seed,
                      int col0, const double* sub, int clamp0, void* vol, int dtype,
                      int ncol, int nx, int ny, int nk, void* stream);

/* ---------------------------------------------------------------------------------------------
 * spcpl.variability_nudge (spcpl.py:613-744, --qt_forcing variance): per (column, level) find beta with
 * mean(max(beta*(qt-<qt>)+<qt>-qsat, 0)) = ql_ref by Brent's method (scipy.optimize.brentq restated),
 * scale the qt fluctuations by it; additive zero-mean noise a*R when beta hits 5; optional constant-T
 * theta_l correction. KJI layout. qt (and thl) are updated in place.
 *   status bits: 1 multiplicative root, 2 nudged to barely unsaturated, 4 additive root,
 *                8 multiplicative bracket failed, 16 additive root not bracketed / R missing,
 *                32 a root search stopped at its iteration limit (scipy raises in cases 16 and 32; the host
 *                mirror sp_coupler_b200/nudge.py raises on them too). */
typedef struct {
  void* qt;              /* in/out [ncol][nk][ny][nx]                                            */
  const void* qsat;      /* [ncol][nk][ny][nx] saturation humidity (les.get_field("Qsat")), or NULL */
  const void* qsat_prof; /* [ncol][nk] horizontally uniform qsat, used when qsat == NULL         */
  void* thl;             /* in/out volume, only with constant_T                                  */
  const void* ql;        /* volume, only with constant_T                                         */
  const double* prof;    /* [5][ncol][nk] slab means from spc_slab_reduce (QT and QL rows)       */
  const void* ql_ref;    /* [ncol][nk] GCM QL on the LES levels (spc_les_forcing.ql_ref)         */
  const void* presf;     /* [ncol][nk] LES pressure, only with constant_T                        */
  const double* R;       /* [ncol][ny][nx] zero-mean normal field (spcpl.py:620-621), may be NULL */
} spc_nudge_io;

int spc_variability_nudge(spc_handle h, const spc_nudge_io* io, int dtype, int ncol, int nx, int ny, int nk,
                          double DT, int constant_T,
                          double* beta /* [ncol][nk] */, double* alpha /* log(beta)/DT */,
                          double* qt_std /* [ncol][nk] */, int32_t* status /* [ncol][nk] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPCPL_B200_H */
