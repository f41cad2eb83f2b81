#include "nccl_dyn.h"

#include <dlfcn.h>
#include <mutex>

namespace devqa {
void set_error(const char* fmt, ...);

static NcclApi g_api;
static bool g_ok = false;
static std::once_flag g_once;

static void load() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // reuse the host application's NCCL
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return;
  g_api.GetVersion = reinterpret_cast<decltype(g_api.GetVersion)>(dlsym(h, "ncclGetVersion"));
  g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  g_api.AllGather = reinterpret_cast<decltype(g_api.AllGather)>(dlsym(h, "ncclAllGather"));
  g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  g_ok = g_api.GetVersion && g_api.GetUniqueId && g_api.CommInitRank && g_api.CommDestroy && g_api.AllGather &&
         g_api.GetErrorString;
}

const NcclApi* nccl_api() {
  std::call_once(g_once, load);
  if (!g_ok) {
    set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
    return nullptr;
  }
  return &g_api;
}

}  // namespace devqa
