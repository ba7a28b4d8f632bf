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

// Optional tensor of a dict: device, contiguity, dtype and element count are all checked, because the C ABI behind
// takes raw pointers and a buffer of the wrong type or length would be written out of bounds on the device.
void* opt(const TDict& d, const char* key, const c10::Device& dev, at::ScalarType st, int64_t numel) {
  if (!d.contains(key)) return nullptr;
  const at::Tensor& t = d.at(key);
  TORCH_CHECK(t.scalar_type() == st, key, " must be ", st, ", got ", t.scalar_type());
  TORCH_CHECK(t.numel() == numel, key, " must have ", numel, " elements, got ", t.numel());
  return const_cast<void*>(cptr(t, dev, key));
}
const void* opt_t(const c10::optional<at::Tensor>& t, const char* name, const c10::Device& dev, at::ScalarType st, int64_t numel) {
  if (!t.has_value()) return nullptr;
  TORCH_CHECK(t->scalar_type() == st, name, " must be ", st, ", got ", t->scalar_type());
  TORCH_CHECK(t->numel() == numel, name, " must have ", numel, " elements, got ", t->numel());
  return cptr(*t, dev, name);
}

const void* req(const TDict& d, const char* key, const c10::Device& dev, at::ScalarType st, int64_t numel) {
  TORCH_CHECK(d.contains(key), "missing GCM/LES field '", key, "'");
  const at::Tensor& t = d.at(key);
  TORCH_CHECK(t.scalar_type() == st, key, " has the wrong dtype");
  TORCH_CHECK(t.numel() == numel, key, " must have ", numel, " elements, got ", t.numel());
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
  const int64_t nf = (int64_t)s.ncol * s.nlev, nh = (int64_t)s.ncol * (s.nlev + 1), nc = s.ncol;
  s.U = req(g, "U", dev, st, nf); s.V = req(g, "V", dev, st, nf); s.T = req(g, "T", dev, st, nf); s.SH = req(g, "SH", dev, st, nf);
  s.QL = req(g, "QL", dev, st, nf); s.QI = req(g, "QI", dev, st, nf); s.Pfull = req(g, "Pfull", dev, st, nf);
  s.A = req(g, "A", dev, st, nf); s.Zgfull = req(g, "Zgfull", dev, st, nf);
  s.Phalf = req(g, "Phalf", dev, st, nh); s.Zghalf = req(g, "Zghalf", dev, st, nh);
  if (surface) {
    s.Z0M = req(g, "Z0M", dev, st, nc); s.Z0H = req(g, "Z0H", dev, st, nc); s.QLflux = req(g, "QLflux", dev, st, nc);
    s.QIflux = req(g, "QIflux", dev, st, nc); s.SHflux = req(g, "SHflux", dev, st, nc); s.TSflux = req(g, "TSflux", dev, st, nc);
    s.TLflux = opt(g, "TLflux", dev, st, nc);
  }
  return s;
}

int64_t mask_words_per_column(int64_t dtype, int64_t layout, int64_t nx, int64_t ny, int64_t nk) {
  return (int64_t)spc_mask_words_per_column((int)dtype, (int)layout, (int)nx, (int)ny, (int)nk);
}

// K1 (spcpl.py:747-759,765)
void slab_reduce(at::TensorList vols, int64_t layout, double ql_thresh, at::Tensor& prof, const c10::optional<at::Tensor>& cnt,
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
  TORCH_CHECK(layout == SPC_LAYOUT_KJI || layout == SPC_LAYOUT_IJK, "bad layout ", layout);
  TORCH_CHECK(prof.scalar_type() == at::kDouble && prof.numel() == (int64_t)5 * ncol * nk, "prof must be float64 [5,ncol,nk]");
  const int64_t mwords = (int64_t)ncol * (int64_t)spc_mask_words_per_column(dtype_code(v0), (int)layout, nx, ny, nk);
  TORCH_CHECK(!mask.has_value() || mwords > 0, "no cloud-mask format for this layout / shape");
  check_rc(spc_slab_reduce(handle_for(dev.index()), p, dtype_code(v0), (int)layout, ncol, nx, ny, nk, ql_thresh,
                           (double*)cptr(prof, dev, "prof"),
                           (int32_t*)opt_t(cnt, "cnt", dev, at::kInt, (int64_t)ncol * nk),
                           (uint32_t*)opt_t(mask, "mask", dev, at::kInt, mwords),
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
  const at::ScalarType st = gcm.at("T").scalar_type();
  const int64_t nk = zf.numel(), ck = (int64_t)g.ncol * nk, cl = (int64_t)g.ncol * g.nlev, nc = g.ncol;
  spc_les_forcing o{};
  o.f_u = opt(out, "f_u", dev, st, ck); o.f_v = opt(out, "f_v", dev, st, ck); o.f_thl = opt(out, "f_thl", dev, st, ck);
  o.f_qt = opt(out, "f_qt", dev, st, ck); o.f_ql = opt(out, "f_ql", dev, st, ck); o.ql_ref = opt(out, "ql_ref", dev, st, ck);
  o.u = opt(out, "u", dev, st, ck); o.v = opt(out, "v", dev, st, ck); o.thl = opt(out, "thl", dev, st, ck);
  o.qt = opt(out, "qt", dev, st, ck);
  o.f_ps = opt(out, "f_ps", dev, st, nc); o.ps = opt(out, "ps", dev, st, nc);
  o.z0m = opt(out, "z0m", dev, st, nc); o.z0h = opt(out, "z0h", dev, st, nc); o.wthl = opt(out, "wthl", dev, st, nc);
  o.wqt = opt(out, "wqt", dev, st, nc);
  o.Tv = opt(out, "Tv", dev, st, cl); o.THL = opt(out, "THL", dev, st, cl); o.QT = opt(out, "QT", dev, st, cl);
  o.Zf = opt(out, "Zf", dev, st, cl);
  o.Zh = opt(out, "Zh", dev, st, cl + nc);
  o.bracket = (int32_t*)opt(out, "bracket", dev, at::kInt, ck);
  o.slab_idx = (int32_t*)opt(out, "slab_idx", dev, at::kInt, cl);
  TORCH_CHECK(!zh.has_value() || (zh->scalar_type() == at::kDouble && zh->numel() == nk), "zh must be float64 [nk]");
  check_rc(spc_gcm_to_les(handle_for(dev.index()), &g, (const double*)cptr(zf, dev, "zf"),
                          zh.has_value() ? (const double*)cptr(*zh, dev, "zh") : nullptr, (int)zf.numel(),
                          (const double*)opt_t(les_prof, "les_prof", dev, at::kDouble, 5 * ck),
                          opt_t(ps_les, "ps_les", dev, st, nc), dt, factor, couple_surface ? 1 : 0,
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
  const at::ScalarType st = gcm.at("T").scalar_type();
  const int64_t nk = zf.numel(), ck = (int64_t)g.ncol * nk, cl = (int64_t)g.ncol * g.nlev, nc = g.ncol;
  TORCH_CHECK(layout == SPC_LAYOUT_KJI || layout == SPC_LAYOUT_IJK, "bad layout ", layout);
  TORCH_CHECK(vol_dtype == SPC_F32 || vol_dtype == SPC_F64, "bad vol_dtype ", vol_dtype);
  TORCH_CHECK(!zh.has_value() || (zh->scalar_type() == at::kDouble && zh->numel() == nk), "zh must be float64 [nk]");
  spc_les_prof l{};
  l.prof = (const double*)req(les, "prof", dev, at::kDouble, 5 * ck);
  l.QL_ice = opt(les, "QL_ice", dev, st, ck); l.T = opt(les, "T", dev, st, ck); l.Rhobf = opt(les, "Rhobf", dev, st, ck);
  l.A = opt(les, "A", dev, st, cl);
  if (les.contains("mask")) {
    TORCH_CHECK(nx > 0 && ny > 0, "mask given without nx, ny");
    const int64_t mwords = nc * (int64_t)spc_mask_words_per_column((int)vol_dtype, (int)layout, (int)nx, (int)ny, (int)nk);
    l.mask = (const uint32_t*)opt(les, "mask", dev, at::kInt, mwords);
  }
  l.slab_idx = (const int32_t*)opt(les, "slab_idx", dev, at::kInt, cl);
  l.cnt = (const int32_t*)opt(les, "cnt", dev, at::kInt, ck);
  l.vol_dtype = (int)vol_dtype; l.layout = (int)layout; l.nx = (int)nx; l.ny = (int)ny;
  spc_gcm_tend o{};
  o.tend = opt(out, "tend", dev, st, cl * SPC_NTEND); o.t = opt(out, "t", dev, st, ck); o.A_d = opt(out, "A_d", dev, st, cl);
  o.cntslab = (int32_t*)opt(out, "cntslab", dev, at::kInt, cl); o.bracket = (int32_t*)opt(out, "bracket", dev, at::kInt, cl);
  o.bracket_pf = (int32_t*)opt(out, "bracket_pf", dev, at::kInt, ck);
  o.start_index = (int32_t*)opt(out, "start_index", dev, at::kInt, nc);
  check_rc(spc_les_to_gcm(handle_for(dev.index()), &g, (const double*)cptr(zf, dev, "zf"),
                          zh.has_value() ? (const double*)cptr(*zh, dev, "zh") : nullptr, (int)zf.numel(), &l, dt, factor,
                          conservative ? 1 : 0, &o, at::cuda::getCurrentCUDAStream(dev.index()).stream()),
           "spc_les_to_gcm");
}

}  // namespace

TORCH_LIBRARY(spcpl_b200, m) {
  m.def("mask_words_per_column(int dtype, int layout, int nx, int ny, int nk) -> int", &mask_words_per_column);
  // outputs are written in place: annotated where the schema language can say so; the `out` dictionaries of the two
  // profile ops are mutated as well (Dict values cannot carry an alias annotation)
  m.def("slab_reduce(Tensor[] vols, int layout, float ql_thresh, Tensor(a!) prof, Tensor(b!)? cnt, Tensor(c!)? mask) -> ()",
        &slab_reduce);
  m.def("gcm_to_les(Dict(str, Tensor) gcm, Tensor zf, Tensor? zh, Tensor? les_prof, Tensor? ps_les, float dt, "
        "float factor, bool couple_surface, Dict(str, Tensor) out) -> ()", &gcm_to_les);
  m.def("les_to_gcm(Dict(str, Tensor) gcm, Tensor zf, Tensor? zh, Dict(str, Tensor) les, int nx, int ny, int layout, "
        "int vol_dtype, float dt, float factor, bool conservative, Dict(str, Tensor) out) -> ()", &les_to_gcm);
}
