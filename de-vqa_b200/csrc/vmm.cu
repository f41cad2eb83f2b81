// Bank storage that grows IN PLACE: CUDA virtual memory management (cuMemAddressReserve + cuMemCreate + cuMemMap).
//
// The lifelong editors append one row per edit (lte_vl.py:150, serac_vl.py:104, recipe_vl.py:163, liveedit.py:223 do it
// with torch.cat -- an O(N) copy per edit).  Round 1 grew the bank geometrically with cudaMalloc + a device-to-device
// copy of everything stored so far: amortised O(1), but every regrow moved up to 30 GB and needed ~2.5x the bank in
// free memory for a moment.  Here the bank lives in one reserved virtual address range and growing it maps fresh
// physical chunks BEHIND the rows already there: no copy, no move, no transient second buffer; the data pointer
// only changes if the reservation itself (4x the first request, at least 4 GiB of address space) is outgrown, and even then the
// physical chunks are re-mapped, not copied.
//
// The driver entry points are fetched with cudaGetDriverEntryPoint (the library does not link libcuda); if they are
// missing or the device refuses (no VMM support), vmm_grow reports it and api.cu keeps the cudaMalloc path.
#include <cuda.h>

#include <new>
#include <vector>

#include "scan.cuh"

namespace devqa {
namespace {

struct DriverApi {
  CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  bool ok = false;
};

template <typename F>
bool entry(const char* name, F* out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(fn);
  return true;
}

const DriverApi& driver() {
  static DriverApi api = [] {
    DriverApi a;
    a.ok = entry("cuMemGetAllocationGranularity", &a.GetGranularity) && entry("cuMemAddressReserve", &a.AddressReserve) &&
           entry("cuMemAddressFree", &a.AddressFree) && entry("cuMemCreate", &a.Create) && entry("cuMemRelease", &a.Release) &&
           entry("cuMemMap", &a.Map) && entry("cuMemUnmap", &a.Unmap) && entry("cuMemSetAccess", &a.SetAccess);
    return a;
  }();
  return api;
}

CUmemAllocationProp props(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}

}  // namespace

struct VmmBuffer {
  int device = 0;
  size_t gran = 0;
  CUdeviceptr base = 0;
  size_t reserved = 0, mapped = 0;
  std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;
};

VmmBuffer* vmm_create(int device) {
  const DriverApi& d = driver();
  if (!d.ok) return nullptr;
  VmmBuffer* b = new (std::nothrow) VmmBuffer();
  if (!b) return nullptr;
  b->device = device;
  const CUmemAllocationProp p = props(device);
  if (d.GetGranularity(&b->gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || b->gran == 0) { delete b; return nullptr; }
  return b;
}

void* vmm_data(const VmmBuffer* b) { return b ? reinterpret_cast<void*>(b->base) : nullptr; }
size_t vmm_mapped(const VmmBuffer* b) { return b ? b->mapped : 0; }

static bool map_all(const DriverApi& d, VmmBuffer* b, CUdeviceptr base) {
  size_t off = 0;
  for (auto& c : b->chunks) {
    if (d.Map(base + off, c.second, 0, c.first, 0) != CUDA_SUCCESS) return false;
    off += c.second;
  }
  if (off) {
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = b->device; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.SetAccess(base, off, &acc, 1) != CUDA_SUCCESS) return false;
  }
  return true;
}

// Make at least want_bytes addressable at vmm_data().  Rows already stored stay where they are (unless the reservation
// is outgrown: then the same physical chunks are mapped into a larger range -- `moved` tells the caller, who has to
// have synchronised work that still uses the old address).  Returns DEVQA_OK, DEVQA_E_NOMEM or DEVQA_E_UNSUPPORTED.
int vmm_grow(VmmBuffer* b, size_t want_bytes, bool* moved) {
  const DriverApi& d = driver();
  if (moved) *moved = false;
  if (!b || !d.ok) return DEVQA_E_UNSUPPORTED;
  if (want_bytes <= b->mapped) return DEVQA_OK;
  const size_t g = b->gran;
  auto round_up = [g](size_t x) { return (x + g - 1) / g * g; };
  // geometric chunks keep the number of handles (and map calls) logarithmic in the final size
  size_t add = round_up(want_bytes - b->mapped);
  const size_t geometric = round_up(b->mapped / 2);
  if (add < geometric) add = geometric;
  const size_t need = b->mapped + add;
  if (need > b->reserved) {
    size_t res = b->reserved ? b->reserved : size_t(4) << 30;
    while (res < need * (b->reserved ? 2 : 4) && res < (size_t(1) << 40)) res <<= 1;
    if (res < need) res = round_up(need);
    CUdeviceptr fresh = 0;
    if (d.AddressReserve(&fresh, res, 0, 0, 0) != CUDA_SUCCESS) {
      set_error("cuMemAddressReserve of %zu bytes failed", res);
      return b->base ? DEVQA_E_NOMEM : DEVQA_E_UNSUPPORTED;
    }
    if (b->base) {
      // re-home the physical chunks: unmap from the old range, map into the new one (no data is copied); nothing on the
      // device may still be using the old addresses
      cudaDeviceSynchronize();
      if (b->mapped) d.Unmap(b->base, b->mapped);
      d.AddressFree(b->base, b->reserved);
      if (!map_all(d, b, fresh)) { set_error("cuMemMap into the enlarged reservation failed"); return DEVQA_E_CUDA; }
      if (moved) *moved = true;
    }
    b->base = fresh;
    b->reserved = res;
  }
  CUmemGenericAllocationHandle h;
  const CUmemAllocationProp p = props(b->device);
  const CUresult cr = d.Create(&h, add, &p, 0);
  if (cr != CUDA_SUCCESS) {
    set_error("cuMemCreate of %zu bytes for the bank failed (CUresult %d)", add, int(cr));
    return (cr == CUDA_ERROR_OUT_OF_MEMORY || b->mapped) ? DEVQA_E_NOMEM : DEVQA_E_UNSUPPORTED;
  }
  if (d.Map(b->base + b->mapped, add, 0, h, 0) != CUDA_SUCCESS) {
    d.Release(h);
    set_error("cuMemMap of %zu bytes failed", add);
    return DEVQA_E_CUDA;
  }
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = b->device; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (d.SetAccess(b->base + b->mapped, add, &acc, 1) != CUDA_SUCCESS) {
    d.Unmap(b->base + b->mapped, add);
    d.Release(h);
    set_error("cuMemSetAccess failed");
    return DEVQA_E_CUDA;
  }
  b->chunks.emplace_back(h, add);
  b->mapped += add;
  return DEVQA_OK;
}

void vmm_destroy(VmmBuffer* b) {
  if (!b) return;
  const DriverApi& d = driver();
  if (d.ok && b->base) {
    if (b->mapped) d.Unmap(b->base, b->mapped);
    for (auto& c : b->chunks) d.Release(c.first);
    d.AddressFree(b->base, b->reserved);
  }
  delete b;
}

}  // namespace devqa
