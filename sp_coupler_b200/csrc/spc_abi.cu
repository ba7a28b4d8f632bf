// Handle management and error reporting of the spcpl_b200 C ABI (include/spcpl_b200.h).
#include <stdarg.h>

#include "spc_common.cuh"

namespace spc {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

int check_handle(spc_handle h) {
  if (h == nullptr || h->magic != SPC_MAGIC) {
    set_error("invalid spc_handle");
    return SPC_ERR_HANDLE;
  }
  return SPC_OK;
}

}  // namespace spc

extern "C" {

int spc_abi_version(void) { return SPC_ABI_VERSION; }

const char* spc_last_error(void) { return spc::g_last_error.c_str(); }

int spc_create(spc_handle* out, int device) {
  SPC_REQUIRE(out != nullptr, SPC_ERR_ARG, "spc_create: out is NULL");
  int ndev = 0;
  SPC_CUDA(cudaGetDeviceCount(&ndev));
  SPC_REQUIRE(device >= 0 && device < ndev, SPC_ERR_ARG, "spc_create: device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  SPC_CUDA(cudaGetDeviceProperties(&prop, device));
  SPC_REQUIRE(prop.major == 10, SPC_ERR_UNSUPPORTED,
              "spc_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
              prop.minor);
  spc_ctx* c = new spc_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->magic = SPC_MAGIC;
  c->k1_variant = c->ijk_variant = c->k2_threads = c->k3_threads = c->proj_threads = 0;
  {
    spc::DeviceGuard guard(device);
    const int rc = spc::k1_configure(c);
    if (rc) {
      delete c;
      return rc;
    }
  }
  *out = c;
  return SPC_OK;
}

int spc_destroy(spc_handle h) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  h->magic = 0;
  delete h;
  return SPC_OK;
}

int spc_host_register(spc_handle h, void* p, size_t nbytes, void** dev_ptr) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(p != nullptr && nbytes > 0 && dev_ptr != nullptr, SPC_ERR_ARG, "spc_host_register: NULL pointer or empty range");
  spc::DeviceGuard guard(h->device);
  SPC_CUDA(cudaHostRegister(p, nbytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
  cudaError_t e = cudaHostGetDevicePointer(dev_ptr, p, 0);
  if (e != cudaSuccess) {
    cudaHostUnregister(p);
    return spc::cuda_fail(e, "cudaHostGetDevicePointer");
  }
  return SPC_OK;
}

int spc_host_unregister(spc_handle h, void* p) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(p != nullptr, SPC_ERR_ARG, "spc_host_unregister: NULL pointer");
  spc::DeviceGuard guard(h->device);
  SPC_CUDA(cudaHostUnregister(p));
  return SPC_OK;
}

int spc_host_device_pointer(spc_handle h, void* p, void** dev_ptr) {
  int rc = spc::check_handle(h);
  if (rc) return rc;
  SPC_REQUIRE(p != nullptr && dev_ptr != nullptr, SPC_ERR_ARG, "spc_host_device_pointer: NULL pointer");
  spc::DeviceGuard guard(h->device);
  SPC_CUDA(cudaHostGetDevicePointer(dev_ptr, p, 0));
  return SPC_OK;
}

}  // extern "C"
