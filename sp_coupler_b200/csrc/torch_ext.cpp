// Thin PyTorch C++ extension over the C ABI (include/spcpl_b200.h): registers the three kernels of
// the coupling step as torch.ops.spcpl_b200.* so that they can be called with tensors from C++ or
// Python without ctypes. It adds no arithmetic: it validates tensors, takes torch's current CUDA
// stream and forwards raw pointers to libspcpl_b200.so. Profile-sized arguments travel as
// Dict(str, Tensor) keyed by the reference's variable names (spcpl.py:32-33).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include <mutex>
#include <unordered_map>

#include "spcpl_b200.h"

namespace {

using TDict = c10::Dict<std::string, at::Tensor>;

spc_handle handle_for(int device) {
  static std::mutex mu;
  static std::unordered_map<int, spc_handle> handles;
  std::lock_guard<std::mutex> lock(mu);
  auto it = handles.find(device);
  if (it != handles.end()) return it->second;
  spc_handle h = nullptr;
  const int rc = spc_create(&h, device);
  TORCH_CHECK(rc == 0, "spc_create failed (", rc, "): ", spc_last_error());
  handles[device] = h;
  return h;
}

int dtype_code(const at::Tensor& t) {
  TORCH_CHECK(t.scalar_type() == at::kFloat || t.scalar_type() == at::kDouble, "expected a float32/float64 tensor");
  return t.scalar_type() == at::kFloat ? SPC_F32 : SPC_F64;
}

const void* cptr(const at::Tensor& t, const c10::Device& dev, const char* name) {
  TORCH_CHECK(t.device() == dev, name, " is on ", t.device(), ", expected ", dev);
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  return t.data_ptr();
}

void* opt(const TDict& d, const char* key, const c10::Device& dev) {
  if (!d.contains(key)) return nullptr;
  return const_cast<void*>(cptr(d.at(key), dev, key));
}

const void* req(const TDict& d, const char* key, const c10::Device& dev, at::ScalarType st) {
  TORCH_CHECK(d.contains(key), "missing GCM/LES field '", key, "'");
  const at::Tensor& t = d.at(key);
  TORCH_CHECK(t.scalar_type() == st, key, " has the wrong dtype");
  return cptr(t, dev, key);
}

void check_rc(int rc, const char* what) { TORCH_CHECK(rc == 0, what, " failed (", rc, "): ", spc_last_error()); }

spc_gcm_cols gcm_struct(const TDict& g, const c10::Device& dev, bool surface) {
  const at::Tensor& T = g.at("T");
  TORCH_CHECK(T.dim() == 2, "gcm['T'] must be [ncol, nlev]");
  const at::ScalarType st = T.scalar_type();
  spc_gcm_cols s{};
  s.ncol = (int)T.size(0);
  s.nlev = (int)T.size(1);
  s.dtype = dtype_code(T);
  s.U = req(g, "U", dev, st); s.V = req(g, "V", dev, st); s.T = req(g, "T", dev, st); s.SH = req(g, "SH", dev, st);
  s.QL = req(g, "QL", dev, st); s.QI = req(g, "QI", dev, st); s.Pfull = req(g, "Pfull", dev, st);
  s.A = req(g, "A", dev, st); s.Zgfull = req(g, "Zgfull", dev, st);
  s.Phalf = req(g, "Phalf", dev, st); s.Zghalf = req(g, "Zghalf", dev, st);
  if (surface) {
    s.Z0M = req(g, "Z0M", dev, st); s.Z0H = req(g, "Z0H", dev, st); s.QLflux = req(g, "QLflux", dev, st);
    s.QIflux = req(g, "QIflux", dev, st); s.SHflux = req(g, "SHflux", dev, st); s.TSflux = req(g, "TSflux", dev, st);
    s.TLflux = opt(g, "TLflux", dev);
  }
  return s;
}

int64_t mask_words_per_column(int64_t dtype, int64_t layout, int64_t nx, int64_t ny, int64_t nk) {
  return (int64_t)spc_mask_words_per_column((int)dtype, (int)layout, (int)nx, (int)ny, (int)nk);
}

// K1 (spcpl.py:747-759,765)
void slab_reduce(at::TensorList vols, int64_t layout, double ql_thresh, at::Tensor prof, const c10::optional<at::Tensor>& cnt,
                 const c10::optional<at::Tensor>& mask) {
  TORCH_CHECK(vols.size() == 5, "slab_reduce expects the five volumes THL, QT, QL, U, V");
  const at::Tensor& v0 = vols[0];
  TORCH_CHECK(v0.is_cuda() && v0.dim() == 4, "volumes must be 4-D CUDA tensors");
  const c10::Device dev = v0.device();
  c10::cuda::CUDAGuard guard(dev);
  const int ncol = (int)v0.size(0);
  const int nk = layout == SPC_LAYOUT_KJI ? (int)v0.size(1) : (int)v0.size(3);
  const int ny = (int)v0.size(2);
  const int nx = layout == SPC_LAYOUT_KJI ? (int)v0.size(3) : (int)v0.size(1);
  const void* p[5];
  for (int f = 0; f < 5; ++f) {
    TORCH_CHECK(vols[f].sizes() == v0.sizes() && vols[f].scalar_type() == v0.scalar_type(), "volume shapes/dtypes differ");
    p[f] = cptr(vols[f], dev, "vol");
  }
  TORCH_CHECK(prof.scalar_type() == at::kDouble && prof.numel() == (int64_t)5 * ncol * nk, "prof must be float64 [5,ncol,nk]");
  check_rc(spc_slab_reduce(handle_for(dev.index()), p, dtype_code(v0), (int)layout, ncol, nx, ny, nk, ql_thresh,
                           (double*)cptr(prof, dev, "prof"),
                           cnt.has_value() ? (int32_t*)cptr(*cnt, dev, "cnt") : nullptr,
                           mask.has_value() ? (uint32_t*)cptr(*mask, dev, "mask") : nullptr,
                           at::cuda::getCurrentCUDAStream(dev.index()).stream()),
           "spc_slab_reduce");
}

// K2 (spcpl.py:171-246, 299-385, 136-167)
void gcm_to_les(const TDict& gcm, const at::Tensor& zf, const c10::optional<at::Tensor>& zh,
                const c10::optional<at::Tensor>& les_prof, const c10::optional<at::Tensor>& ps_les, double dt, double factor,
                bool couple_surface, const TDict& out) {
  const c10::Device dev = zf.device();
  TORCH_CHECK(dev.is_cuda() && zf.scalar_type() == at::kDouble, "zf must be a float64 CUDA tensor");
  c10::cuda::CUDAGuard guard(dev);
  spc_gcm_cols g = gcm_struct(gcm, dev, couple_surface);
  spc_les_forcing o{};
  o.f_u = opt(out, "f_u", dev); o.f_v = opt(out, "f_v", dev); o.f_thl = opt(out, "f_thl", dev);
  o.f_qt = opt(out, "f_qt", dev); o.f_ql = opt(out, "f_ql", dev); o.ql_ref = opt(out, "ql_ref", dev);
  o.u = opt(out, "u", dev); o.v = opt(out, "v", dev); o.thl = opt(out, "thl", dev); o.qt = opt(out, "qt", dev);
  o.f_ps = opt(out, "f_ps", dev); o.ps = opt(out, "ps", dev);
  o.z0m = opt(out, "z0m", dev); o.z0h = opt(out, "z0h", dev); o.wthl = opt(out, "wthl", dev); o.wqt = opt(out, "wqt", dev);
  o.Tv = opt(out, "Tv", dev); o.THL = opt(out, "THL", dev); o.QT = opt(out, "QT", dev); o.Zf = opt(out, "Zf", dev);
  o.Zh = opt(out, "Zh", dev);
  o.bracket = (int32_t*)opt(out, "bracket", dev);
  o.slab_idx = (int32_t*)opt(out, "slab_idx", dev);
  check_rc(spc_gcm_to_les(handle_for(dev.index()), &g, (const double*)cptr(zf, dev, "zf"),
                          zh.has_value() ? (const double*)cptr(*zh, dev, "zh") : nullptr, (int)zf.numel(),
                          les_prof.has_value() ? (const double*)cptr(*les_prof, dev, "les_prof") : nullptr,
                          ps_les.has_value() ? cptr(*ps_les, dev, "ps_les") : nullptr, dt, factor, couple_surface ? 1 : 0,
                          &o, at::cuda::getCurrentCUDAStream(dev.index()).stream()),
           "spc_gcm_to_les");
}

// K3 (spcpl.py:388-555)
void les_to_gcm(const TDict& gcm, const at::Tensor& zf, const c10::optional<at::Tensor>& zh, const TDict& les, int64_t nx,
                int64_t ny, int64_t layout, int64_t vol_dtype, double dt, double factor, bool conservative, const TDict& out) {
  const c10::Device dev = zf.device();
  TORCH_CHECK(dev.is_cuda() && zf.scalar_type() == at::kDouble, "zf must be a float64 CUDA tensor");
  c10::cuda::CUDAGuard guard(dev);
  spc_gcm_cols g = gcm_struct(gcm, dev, false);
  spc_les_prof l{};
  l.prof = (const double*)req(les, "prof", dev, at::kDouble);
  l.QL_ice = opt(les, "QL_ice", dev); l.T = opt(les, "T", dev); l.Rhobf = opt(les, "Rhobf", dev); l.A = opt(les, "A", dev);
  l.mask = (const uint32_t*)opt(les, "mask", dev);
  l.slab_idx = (const int32_t*)opt(les, "slab_idx", dev);
  l.cnt = (const int32_t*)opt(les, "cnt", dev);
  l.vol_dtype = (int)vol_dtype; l.layout = (int)layout; l.nx = (int)nx; l.ny = (int)ny;
  spc_gcm_tend o{};
  o.tend = opt(out, "tend", dev); o.t = opt(out, "t", dev); o.A_d = opt(out, "A_d", dev);
  o.cntslab = (int32_t*)opt(out, "cntslab", dev); o.bracket = (int32_t*)opt(out, "bracket", dev);
  o.bracket_pf = (int32_t*)opt(out, "bracket_pf", dev); o.start_index = (int32_t*)opt(out, "start_index", dev);
  check_rc(spc_les_to_gcm(handle_for(dev.index()), &g, (const double*)cptr(zf, dev, "zf"),
                          zh.has_value() ? (const double*)cptr(*zh, dev, "zh") : nullptr, (int)zf.numel(), &l, dt, factor,
                          conservative ? 1 : 0, &o, at::cuda::getCurrentCUDAStream(dev.index()).stream()),
           "spc_les_to_gcm");
}

}  // namespace

TORCH_LIBRARY(spcpl_b200, m) {
  m.def("mask_words_per_column(int dtype, int layout, int nx, int ny, int nk) -> int", &mask_words_per_column);
  m.def("slab_reduce(Tensor[] vols, int layout, float ql_thresh, Tensor prof, Tensor? cnt, Tensor? mask) -> ()", &slab_reduce);
  m.def("gcm_to_les(Dict(str, Tensor) gcm, Tensor zf, Tensor? zh, Tensor? les_prof, Tensor? ps_les, float dt, "
        "float factor, bool couple_surface, Dict(str, Tensor) out) -> ()", &gcm_to_les);
  m.def("les_to_gcm(Dict(str, Tensor) gcm, Tensor zf, Tensor? zh, Dict(str, Tensor) les, int nx, int ny, int layout, "
        "int vol_dtype, float dt, float factor, bool conservative, Dict(str, Tensor) out) -> ()", &les_to_gcm);
}
