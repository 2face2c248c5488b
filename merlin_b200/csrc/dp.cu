// merlin_b200 — data-parallel gradient exchange inside the library (merlin_dp_*): one NCCL communicator per
// rank over NVLink 5 / NVSwitch, a bucketed in-place bf16 all-reduce on a caller-given stream, and a CTA
// budget so the exchange co-resides with the persistent tcgen05 GEMMs of backward instead of waiting for
// whole SMs (see merlin_set_sm_budget in gemm.cu).  Replaces the FSDP reduce-scatter / all-gather of the
// reference's launch line (playground/merlin/clip-large+conv+vicuna-v15-7b/pretrain.sh:32-33).
//
// libnccl is resolved at run time, in this order: the path given to merlin_dp_load_nccl() / $MERLIN_NCCL_LIB;
// the libnccl.so.2 the process has already loaded (RTLD_NOLOAD — under PyTorch that is the torch-bundled one;
// loading a second, different libnccl.so.2 by bare name BEFORE torch would make the loader hand torch the
// wrong SONAME match); finally the loader's search path (plain C/C++ hosts).  Only prototypes from <nccl.h>
// are needed at build time.
#include "host.cuh"

#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>

namespace mb {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommRegister)(const ncclComm_t, void*, size_t, void**) = nullptr;
  ncclResult_t (*CommDeregister)(const ncclComm_t, void*) = nullptr;
  ncclResult_t (*MemAlloc)(void**, size_t) = nullptr;
  ncclResult_t (*MemFree)(void*) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  const char* (*GetLastError)(ncclComm_t) = nullptr;
  bool ok = false;
};

char g_nccl_path[1024] = "";

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    const char* env = getenv("MERLIN_NCCL_LIB");
    if (g_nccl_path[0]) a.handle = dlopen(g_nccl_path, RTLD_NOW | RTLD_LOCAL);
    if (!a.handle && env && env[0]) a.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!a.handle)
      for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
        if (a.handle) break;
      }
    if (!a.handle)
      for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (a.handle) break;
      }
    if (!a.handle) return;
#define MB_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym))
    MB_SYM(GetUniqueId, "ncclGetUniqueId");
    MB_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
    MB_SYM(AllReduce, "ncclAllReduce");
    MB_SYM(CommDestroy, "ncclCommDestroy");
    MB_SYM(CommRegister, "ncclCommRegister");
    MB_SYM(CommDeregister, "ncclCommDeregister");
    MB_SYM(MemAlloc, "ncclMemAlloc");
    MB_SYM(MemFree, "ncclMemFree");
    MB_SYM(GetVersion, "ncclGetVersion");
    MB_SYM(GetErrorString, "ncclGetErrorString");
    MB_SYM(GetLastError, "ncclGetLastError");
#undef MB_SYM
    a.ok = a.GetUniqueId && a.CommInitRankConfig && a.AllReduce && a.CommDestroy && a.GetErrorString;
  });
  return a;
}

struct DpComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, max_ctas = 0;
};

}  // namespace
}  // namespace mb

#define MB_CHECK_NCCL(expr)                                                                   \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      ::mb::set_error("%s failed: %s (%s:%d)", #expr, ::mb::api().GetErrorString(_r), __FILE__, __LINE__); \
      return MERLIN_E_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define MB_NEED_NCCL()                                                                        \
  do {                                                                                        \
    if (!::mb::api().ok) {                                                                    \
      ::mb::set_error("libnccl.so.2 could not be resolved (dlopen/dlsym): %s", dlerror());    \
      return MERLIN_E_UNSUPPORTED;                                                            \
    }                                                                                         \
  } while (0)

extern "C" {

int merlin_dp_load_nccl(const char* path) {
  if (path && path[0]) snprintf(mb::g_nccl_path, sizeof(mb::g_nccl_path), "%s", path);
  MB_NEED_NCCL();
  return MERLIN_OK;
}

int merlin_dp_unique_id(void* id_out) {
  MB_NEED_NCCL();
  MB_CHECK_ARG(id_out != nullptr, "merlin_dp_unique_id: null output");
  static_assert(sizeof(ncclUniqueId) == MERLIN_DP_UNIQUE_ID_BYTES, "ncclUniqueId size");
  MB_CHECK_NCCL(mb::api().GetUniqueId(static_cast<ncclUniqueId*>(id_out)));
  return MERLIN_OK;
}

int merlin_dp_init(const void* unique_id, int rank, int world, int max_ctas, void** comm_out) {
  MB_NEED_NCCL();
  MB_CHECK_ARG(unique_id && comm_out && world >= 1 && rank >= 0 && rank < world,
               "merlin_dp_init: bad arguments (rank %d of %d)", rank, world);
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
  cfg.blocking = 1;
  if (max_ctas > 0) {
    cfg.minCTAs = 1;
    cfg.maxCTAs = max_ctas;
  }
  auto* c = new mb::DpComm;
  c->rank = rank;
  c->world = world;
  c->max_ctas = max_ctas;
  ncclResult_t r = mb::api().CommInitRankConfig(&c->comm, world, id, rank, &cfg);
  if (r != ncclSuccess) {
    mb::set_error("ncclCommInitRankConfig failed: %s", mb::api().GetErrorString(r));
    delete c;
    return MERLIN_E_CUDA;
  }
  *comm_out = c;
  return MERLIN_OK;
}

int merlin_dp_allreduce_bucket(void* comm, void* stream, void* grads, int64_t count) {
  MB_NEED_NCCL();
  auto* c = static_cast<mb::DpComm*>(comm);
  MB_CHECK_ARG(c && c->comm && grads && count >= 0, "merlin_dp_allreduce_bucket: bad arguments");
  if (count == 0 || c->world == 1) return MERLIN_OK;
  MB_CHECK_NCCL(mb::api().AllReduce(grads, grads, static_cast<size_t>(count), ncclBfloat16, ncclSum, c->comm,
                                    static_cast<cudaStream_t>(stream)));
  return MERLIN_OK;
}

int merlin_dp_allreduce_f32(void* comm, void* stream, void* buf, int64_t count, int op_max) {
  MB_NEED_NCCL();
  auto* c = static_cast<mb::DpComm*>(comm);
  MB_CHECK_ARG(c && c->comm && buf && count >= 0, "merlin_dp_allreduce_f32: bad arguments");
  if (count == 0 || c->world == 1) return MERLIN_OK;
  MB_CHECK_NCCL(mb::api().AllReduce(buf, buf, static_cast<size_t>(count), ncclFloat32, op_max ? ncclMax : ncclSum,
                                    c->comm, static_cast<cudaStream_t>(stream)));
  return MERLIN_OK;
}

int merlin_dp_register(void* comm, void* buf, int64_t bytes, void** handle_out) {
  MB_NEED_NCCL();
  auto* c = static_cast<mb::DpComm*>(comm);
  MB_CHECK_ARG(c && c->comm && buf && bytes > 0 && handle_out, "merlin_dp_register: bad arguments");
  MB_CHECK_ARG(mb::api().CommRegister != nullptr, "merlin_dp_register: this libnccl has no ncclCommRegister");
  MB_CHECK_NCCL(mb::api().CommRegister(c->comm, buf, static_cast<size_t>(bytes), handle_out));
  return MERLIN_OK;
}

int merlin_dp_deregister(void* comm, void* handle) {
  MB_NEED_NCCL();
  auto* c = static_cast<mb::DpComm*>(comm);
  MB_CHECK_ARG(c && c->comm && handle, "merlin_dp_deregister: bad arguments");
  MB_CHECK_NCCL(mb::api().CommDeregister(c->comm, handle));
  return MERLIN_OK;
}

int merlin_dp_mem_alloc(void** ptr_out, int64_t bytes) {
  MB_NEED_NCCL();
  MB_CHECK_ARG(ptr_out && bytes > 0, "merlin_dp_mem_alloc: bad arguments");
  MB_CHECK_ARG(mb::api().MemAlloc != nullptr, "merlin_dp_mem_alloc: this libnccl has no ncclMemAlloc");
  MB_CHECK_NCCL(mb::api().MemAlloc(ptr_out, static_cast<size_t>(bytes)));
  return MERLIN_OK;
}

int merlin_dp_mem_free(void* ptr) {
  MB_NEED_NCCL();
  if (ptr == nullptr) return MERLIN_OK;
  MB_CHECK_NCCL(mb::api().MemFree(ptr));
  return MERLIN_OK;
}

int merlin_dp_finalize(void* comm) {
  auto* c = static_cast<mb::DpComm*>(comm);
  if (c == nullptr) return MERLIN_OK;
  MB_NEED_NCCL();
  ncclResult_t r = c->comm ? mb::api().CommDestroy(c->comm) : ncclSuccess;
  delete c;
  if (r != ncclSuccess) {
    mb::set_error("ncclCommDestroy failed: %s", mb::api().GetErrorString(r));
    return MERLIN_E_CUDA;
  }
  return MERLIN_OK;
}

int merlin_dp_nccl_version(void) {
  if (!mb::api().ok || !mb::api().GetVersion) return 0;
  int v = 0;
  mb::api().GetVersion(&v);
  return v;
}

}  // extern "C"
